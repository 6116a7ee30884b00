"""Host <-> device plumbing shared by the reference-signature wrappers."""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mrflow_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def up(a, dtype):
    """numpy -> contiguous CUDA tensor of ``dtype`` (numpy casting semantics of ``astype``)."""
    a = np.ascontiguousarray(np.asarray(a), dtype=dtype)
    return torch.from_numpy(a).cuda(non_blocking=True)


def down(t):
    """device -> numpy.  Anything sizeable goes through a pinned staging buffer (55 GB/s on this box
    instead of ~10 GB/s pageable); the returned array owns that buffer."""
    if t.numel() * t.element_size() < (1 << 20):
        return t.cpu().numpy()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def flag(a):
    """boolean-ish array -> uint8 CUDA tensor"""
    return up(np.asarray(a).astype(bool), np.uint8)
