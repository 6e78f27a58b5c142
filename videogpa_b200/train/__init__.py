"""Training entry points mirroring the reference's `train/*/03_train.py` scripts on the sm_100a kernels."""
