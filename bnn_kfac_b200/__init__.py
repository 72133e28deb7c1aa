"""bnn_kfac_b200 — B200-native Kronecker-factored Laplace engine.

Drop-in for the hot path of TianmingQiu/BNN_KFAC (`models/curvatures.py` KFAC / Diagonal and the
sampling / prediction loops built on it), backed by hand-written sm_100a kernels in libbk_kfac.so.
"""
__version__ = "0.1.0"
