"""The C++ host side: drop-in templates (include/huygens/*.h) and the offline render harness, built
with g++ against libhuygens_cuda.so, checked against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import rel_err
from huygens_b200 import design

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "harness")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    return os.path.join(ROOT, "harness")


def _resynthesis_oracle(N, dtype, x):
    freqs, fwd, back = design.resynthesis_bank(N)
    fb = oracle.Filterbank(2, N, 0.1, 1.0, dtype=dtype)
    for n in range(N):
        fb.coefficients(n, fwd[n], back[n])
    fb.boost(np.ones(N))
    fb.open()
    return fb.process(x.astype(dtype))


@pytest.mark.parametrize("flags,dtype", [(["--double"], np.float64), ([], np.float32),
                                         (["--double", "--per-sample", "--bsize", "64"], np.float64)])
def test_render_harness_matches_oracle(harness, tmp_path, flags, dtype, oracle_lib):
    N, seconds = 64, 0.25 if "--per-sample" not in flags else 0.02
    out = tmp_path / "out.f32"
    r = subprocess.run([os.path.join(harness, "render"), "--filters", str(N), "--seconds", str(seconds),
                        "--input", "white", "--out", str(out)] + flags, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    L = int(seconds * 48000)
    assert info["frames"] == L and info["filters"] == N
    got = np.fromfile(out, dtype=np.float32)
    x = design.white_noise(L, dtype=np.float32)          # the harness generator must be the same stream
    ref = _resynthesis_oracle(N, dtype, x)
    assert rel_err(got, ref.astype(np.float32)) <= 2e-6


def test_render_harness_reads_and_writes_audio_files(harness, tmp_path, oracle_lib):
    """the harness renders from a WAV (PCM16) or raw float32 file and writes float32 WAV."""
    import struct
    import wave
    N, L = 32, 6000
    x16 = (design.white_noise(L) * 20000).astype(np.int16)
    with wave.open(str(tmp_path / "in.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(48000); w.writeframes(x16.tobytes())
    out = tmp_path / "out.wav"
    r = subprocess.run([os.path.join(harness, "render"), "--filters", str(N), "--double", "--in", str(tmp_path / "in.wav"),
                        "--out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE" and struct.unpack("<H", raw[20:22])[0] == 3      # IEEE float
    got = np.frombuffer(raw[44:], dtype=np.float32)
    x = (x16.astype(np.float32) / 32768.0)
    ref = _resynthesis_oracle(N, np.float64, x)
    assert len(got) == L and rel_err(got, ref.astype(np.float32)) <= 2e-6
    x.tofile(tmp_path / "in.f32")
    r = subprocess.run([os.path.join(harness, "render"), "--filters", str(N), "--double", "--in", str(tmp_path / "in.f32"),
                        "--out", str(tmp_path / "out.f32")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "out.f32", dtype=np.float32), got)


def test_render_harness_c1_full(harness, tmp_path, oracle_lib):
    """BASELINE configs[0]: Filterbank<double>, 64 resonant band-passes, 10 s of 48 kHz white noise."""
    out = tmp_path / "c1.f32"
    r = subprocess.run([os.path.join(harness, "render"), "--filters", "64", "--seconds", "10", "--double",
                        "--bsize", "1024", "--out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out, dtype=np.float32)
    ref = _resynthesis_oracle(64, np.float64, design.white_noise(480000, dtype=np.float32))
    assert rel_err(got, ref.astype(np.float32)) <= 2e-6


def test_dropin_per_sample_banks(harness, oracle_lib):
    ticks = 80
    r = subprocess.run([os.path.join(harness, "banks_demo"), str(ticks)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    extra = {ln.split()[1]: float(ln.split()[2]) for ln in lines if ln.startswith("#")}
    got = np.array([[float(v) for v in line.split()] for line in lines if not line.startswith("#")])
    assert got.shape == (ticks, 10)
    # the block entry points of the drop-in header: fused followers (replica mode) and a parameter schedule
    NB, L, width, hop = 16, 1200, 96, 32

    def mk():
        fb = oracle.Filterbank(2, NB, 0.01, 0.05)
        for i in range(NB):
            fb.coefficients(i, [1.0, 0.0, -1.0], [-1.9 + 0.05 * i, 0.98])
        fb.open()
        fb.boost(np.ones(NB))
        return fb
    x = design.white_noise(L)
    ys, bands = mk().process(x, bands=True)
    env = oracle.RMSbank(NB, width).process(bands)[:, hop - 1::hop][:, : L // hop]
    assert abs(extra["env_sum"] - env.sum()) <= 1e-12 * abs(env.sum())
    assert abs(extra["env_out_sum"] - ys.sum()) <= 1e-12 * np.abs(ys).sum()
    fb = mk()
    y = [fb.process(x[:300])]
    fb.mix(2, 0.5)
    y.append(fb.process(x[300:700]))
    fb.boost(5, 0.5)
    y.append(fb.process(x[700:]))
    e = float((np.concatenate(y) ** 2).sum())
    assert abs(extra["events_energy"] - e) <= 1e-12 * e
    osc = oracle.Oscbank(8)
    for i in range(8):
        osc.freqmod(i, 110.0 * (i + 1))
    osc.open()
    osc.deactivate([3])
    mix, ph = osc.process(ticks, phases=True)
    assert np.array_equal(got[:, 0], mix.real) and np.array_equal(got[:, 1], mix.imag)
    zre, zim = np.ascontiguousarray(ph.real.T), np.ascontiguousarray(ph.imag.T)
    first = oracle.RMSbank(8, 16).process(zre, zim)
    again = oracle.RMSbank(8, 16).process(zre, zim, calls_per_tick=2)
    assert np.array_equal(got[:, 2], first[0]) and np.array_equal(got[:, 3], again[0])
    gated = oracle.Latchbank(8, 0.9, 0.2, width=16).process(zre, zim)[0]
    assert np.allclose(got[:, 6], gated.sum(axis=0), rtol=0, atol=1e-13)
    weighted = (ph.real * ((np.arange(8) + 1) / 8.0)).sum(axis=1)
    assert np.allclose(got[:, 9], weighted, rtol=0, atol=1e-13)
    radii = np.array([complex(0.99 - 0.01 * i, 0.002 * i) for i in range(8)])
    smoothed = oracle.Stickbank(8, 1, -0.9).process(oracle.Slidebank(8, 2, radii).process(ph.T))
    assert np.allclose(got[:, 7] + 1j * got[:, 8], smoothed.sum(axis=0), rtol=0, atol=1e-13)
    F = oracle.Filterbank(2, 4, 0.001, 1.0, fixed=True)
    for i in range(4):
        F.coefficients(i, [0.5, 0, -0.5], [-1.6 + 0.1 * i, 0.81])
    F.open()
    F.boost(np.ones(4))
    for t in range(ticks):
        y, bands = F.process(mix.real[t:t + 1], bands=True)
        _, g = F.smoothed()
        assert abs(got[t, 4] - y[0]) <= 1e-12 * max(1.0, abs(y[0]))
        clipped = np.clip(bands[:, 0] * g, -0.125, 0.125).sum()
        assert abs(got[t, 5] - clipped) <= 1e-12 * max(1.0, abs(clipped))


def test_sharded_demo_single_process_all_gpus(harness, tmp_path, oracle_lib):
    """harness/sharded_demo: one C++ host process, hg_comm_init_all, one soundmath::Filterbank shard per GPU, the
    running bank-sum row handed GPU to GPU (every GPU of the box; a one-GPU box runs the same code with one rank).
    The first block must be the oracle's single-bank sum, bit for bit."""
    import torch
    from conftest import oracle_chained
    G = max(1, min(torch.cuda.device_count(), 4))
    N, seconds = 600 * G + 7, 0.25
    out = tmp_path / "sum.f32"
    r = subprocess.run([os.path.join(harness, "sharded_demo"), "--filters", str(N), "--seconds", str(seconds), "--gpus", str(G),
                        "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    L = int(seconds * 48000)
    assert info["gpus"] == G and info["frames"] == L and info["exchange"] == "chain"
    got = np.fromfile(out, dtype=np.float32)
    ref = oracle_chained(N, np.float32, design.chirp(L, dtype=np.float32))
    assert np.array_equal(got, ref)
    r = subprocess.run([os.path.join(harness, "sharded_demo"), "--filters", str(N), "--seconds", str(seconds), "--gpus", str(G),
                        "--reduce", "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert rel_err(np.fromfile(out, dtype=np.float32), ref) <= 1e-5
