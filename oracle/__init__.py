"""CPU oracle for the preset-gen-vae hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain CPU PyTorch / NumPy, the algorithms of the
reference hot path (utils/audio.py front end, model/{encoder,decoder,layer,VAE,
flows,regression,loss}.py, utils/probability.py) and of the two third-party
dependencies that are absent from /root/reference (nflows ~=0.14, librosa
~=0.8.0).  Nothing in the product path (preset_gen_vae_b200/) imports it; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs do, and only as the checker or as the timed CPU baseline.

Parity status
-------------
* Pinned against the reference itself: every module that /root/reference ships
  (front end STFT/dB, encoder, decoder, VAE wrappers, regression head, losses,
  probability helpers, preset index tables) is validated by
  oracle/make_golden.py, which imports the reference read-only, runs it on the
  seeded synthetic inputs and (a) asserts the restatement matches, (b) writes
  tests/golden/*.npz from the REFERENCE's outputs.
* PARITY UNPINNED: the nflows classes (oracle/nflows_port.py) and the librosa
  mel filterbank (oracle/frontend.py::slaney_mel_filterbank) are restated from
  the published algorithms because neither package is installed here and the
  reference holds no test or golden vector for them.  The mel matrix is
  cross-checked against torchaudio's independent Slaney implementation.
"""
