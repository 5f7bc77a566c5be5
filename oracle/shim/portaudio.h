// empty stand-in: /root/reference/src/includes.h:18 includes <portaudio.h>, the bank headers never use it.
