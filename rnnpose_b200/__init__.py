"""rnnpose_b200 -- Blackwell-native recurrent pose-refinement inner loop (RNNPose hot path).

Only what the hot path needs lives here: ``csrc/`` (sm_100a CUDA kernels + the
C-ABI library ``libb200pose.so``), the ctypes loader, the host-side mirror of the reference's
``PoseRefiner`` / ``SE3Sequence`` interface, the synthetic-scene generator and the shard/all-gather
helper.  There is no CPU fallback: importing :mod:`rnnpose_b200.ops` works without a GPU (so the
symbol-export test can run), but every compute entry point raises if the library or a CUDA device
is missing.
"""
__version__ = "0.1.0"
