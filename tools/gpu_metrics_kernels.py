"""Launches the calibration-metric kernels once per shape (for `ncu --metrics gpu__time_duration.sum,...`)."""
import sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import utilities as U
dev = torch.device("cuda:0")
for rows, classes in ((1 << 22, 10), (1 << 17, 1000)):
    p = torch.softmax(torch.randn(rows, classes, device=dev), 1)
    lab = torch.randint(0, classes, (rows,), device=dev)
    for _ in range(3):
        U.expected_calibration_error(p, lab)
torch.cuda.synchronize()
