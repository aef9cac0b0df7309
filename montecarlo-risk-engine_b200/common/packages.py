"""Host-side globals (reference: src/common/packages.py:1-11).

``device`` is the device of the small host-side tensors the API exposes
(timelines, model parameters).  The Monte Carlo itself always runs on the
CUDA device of this process (``mcre.runtime.compute_device()``); there is no
CPU fallback for it.
"""
import torch

device = torch.device("cpu")
FLOAT = torch.float64
