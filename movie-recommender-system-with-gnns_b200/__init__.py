"""B200-native LightGCN hot path (sm_100a CUDA behind a C-ABI, PyTorch host code).

Module layout mirrors the reference (/root/reference): ``models.light_gcn``,
``utils.helpers``, ``utils.train_test``, ``utils.recommend``, ``data.dataset_handler``.
There is no CPU fallback: every op raises if ``csrc/liblgcn_b200.so`` is missing or the
tensors are not on a CUDA device.
"""
__version__ = "0.1.0"
