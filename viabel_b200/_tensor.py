"""Host<->device plumbing: the public API accepts numpy arrays (as the reference does) or
CUDA tensors; results come back in the kind that went in."""
import numpy as np
import torch

F64 = torch.float64


def device():
    if not torch.cuda.is_available():
        raise RuntimeError('viabel_b200 needs a CUDA device (sm_100a); there is no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def is_host(x):
    return not isinstance(x, torch.Tensor)


def to_dev(x, dtype=F64):
    """numpy / list / scalar / tensor -> contiguous CUDA tensor of `dtype`."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device(), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)), dtype=dtype).to(device())


def like_input(t, ref_was_host):
    """Return numpy if the user gave numpy, else the tensor itself."""
    if ref_was_host:
        a = t.detach().cpu().numpy()
        return a if a.ndim else a[()]
    return t
