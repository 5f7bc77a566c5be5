// empty stand-in: /root/reference/src/includes.h:25 includes "RtMidi.h", the bank headers never use it.
