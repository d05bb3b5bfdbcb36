"""marconet_b200 -- B200-native (sm_100a) implementation of the MARCONet inference hot path.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/marconet_b200.h)
  _lib.py    ctypes binding of libmarconet_b200.so (fails loudly when missing)
  ops.py     torch-tensor wrappers over the C ABI (device pointers + current stream)
  models/    host-side mirror of the reference's models/{networks,resnet,textvit_arch}.py API
  parallel.py  line / character sharding over torch.distributed (NCCL)
"""
__version__ = "0.1.0"
