"""B200-native hot path of gwendal-lv/preset-gen-vae (spectrogram front end, conv VAE, RealNVP flows, preset
regression head, losses) behind the reference's model/build.py + ExtendedAE + config.py API.

All device work is done by hand-written sm_100a CUDA kernels in `csrc/`, reached through the C ABI declared in
`include/pgv.h` (`libpgv.so`, loaded by `_lib.py`).  There is no CPU or eager-PyTorch fallback: importing a module
that needs the extension raises if `libpgv.so` is missing.
"""
__version__ = "0.1.0"
