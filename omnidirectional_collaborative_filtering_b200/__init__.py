"""B200-native training/scoring hot path for the omnidirectional-CF autoencoder family.

Host side: Python mirrors of the reference's `data_reader` / `omni_model` / `train.py`
surface. Device side: hand-written sm_100a CUDA behind the C ABI in `include/ocf.h`
(`csrc/libocf_b200.so`). There is no CPU fallback: anything that computes raises if the
library cannot be loaded.
"""
__version__ = "0.1.0"
