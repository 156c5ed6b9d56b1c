"""PF8 activation buffers: the HBM layout every conv of the path reads and writes.

Logical [N, C, H, W]  ->  C/8 planes x P positions x 8 bf16, P = N*(H+1)*(W+1), position
p = (n*(H+1) + y+1)*(W+1) + x+1 ; row 0 / column 0 of every image are shared zero padding.
Each plane carries zero guard bands so the conv's halo loads never leave the allocation.
"""
import ctypes as C
import os

import torch

from . import _lib

_CANARY = int(os.environ.get("HRNB_CANARY", "0"))    # debug: NaN margins (elements) around every PF8 allocation


class _Arena:
    """Zero-filled device memory handed out in 256-byte-aligned slices of large chunks: a launch plan owns several hundred
    PF8 tensors, and one fill kernel per tensor made plan construction a stream of ATen launches.  A chunk is released by
    the caching allocator when the last tensor carved from it dies (slices keep the chunk's storage alive)."""
    CHUNK = 256 << 20

    def __init__(self):
        self.cur = {}        # device -> [chunk tensor (uint8), bytes used]

    def take(self, nbytes, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        nbytes = (nbytes + 255) // 256 * 256
        if nbytes >= self.CHUNK // 4:
            return torch.zeros(nbytes, dtype=torch.uint8, device=device)
        slot = self.cur.get(device)
        if slot is None or slot[1] + nbytes > slot[0].numel():
            slot = self.cur[device] = [torch.zeros(self.CHUNK, dtype=torch.uint8, device=device), 0]
        out = slot[0][slot[1]:slot[1] + nbytes]
        slot[1] += nbytes
        return out


_ARENA = _Arena()


def _alloc(shape, device):
    """zero-filled bf16 buffer (a slice of the arena); with HRNB_CANARY=n it sits between two n-element NaN margins so that
    any kernel reading outside a PF8 allocation produces NaNs deterministically (debug aid)"""
    n = 1
    for d in shape:
        n *= d
    if not _CANARY:
        return _ARENA.take(n * 2, device)[:n * 2].view(torch.bfloat16).view(shape)
    raw = torch.full((n + 2 * _CANARY,), float("nan"), dtype=torch.bfloat16, device=device)
    body = raw[_CANARY:_CANARY + n]
    body.zero_()
    return body.view(shape)


class PF8:
    __slots__ = ("buf", "N", "C", "H", "W", "Hp", "Wp", "P", "planes", "ps", "lead", "_ptr")

    def __init__(self, N, C_, H, W, device="cuda", buf=None, plane_offset=0, planes_total=None):
        assert C_ % 8 == 0, "PF8 needs channels % 8 == 0"
        self.N, self.C, self.H, self.W = N, C_, H, W
        self.Hp, self.Wp = H + 1, W + 1
        self.P = N * self.Hp * self.Wp
        self.planes = C_ // 8
        self.lead = _lib.guard_lead(self.Wp)
        tail = _lib.guard_tail(self.Wp)
        self.ps = self.lead + (self.P + 7) // 8 * 8 + tail
        if buf is None:
            buf = _alloc((planes_total or self.planes, self.ps, 8), device)
        self.buf = buf
        self._ptr = buf.data_ptr() + (plane_offset * self.ps + self.lead) * 16

    @property
    def ptr(self):
        """device address of position 0 of plane 0"""
        return self._ptr

    def view_planes(self, plane0, nplanes):
        """A PF8 tensor aliasing planes [plane0, plane0+nplanes) of this buffer (used for the concat)."""
        v = PF8.__new__(PF8)
        for s in ("N", "H", "W", "Hp", "Wp", "P", "ps", "lead", "buf"):
            setattr(v, s, getattr(self, s))
        v.C = nplanes * 8
        v.planes = nplanes
        v._ptr = self._ptr + plane0 * self.ps * 16
        return v

    # ---- conversions through the CUDA kernels --------------------------------------------------
    def to_nchw(self):
        out = torch.empty((self.N, self.C, self.H, self.W), dtype=torch.float32, device=self.buf.device)
        _lib.check(_lib.lib().hrnb_pf8_to_nchw_f32(self.ptr, self.ps, self.N, self.C, self.H, self.W,
                                                   out.data_ptr(), _lib.stream_ptr()))
        return out

    @staticmethod
    def from_nchw(x):
        x = x.contiguous().float()
        N, C_, H, W = x.shape
        Cp = (C_ + 7) // 8 * 8
        t = PF8(N, Cp, H, W, device=x.device)
        _lib.check(_lib.lib().hrnb_nchw_f32_to_pf8(x.data_ptr(), N, C_, H, W, t.ptr, t.ps, _lib.stream_ptr()))
        return t

    # ---- pure-torch views (tests only; no kernels) ------------------------------------------------
    def torch_interior(self):
        """[N, C, H, W] float32 gathered with torch indexing (test helper)."""
        body = self.buf[:, self.lead:self.lead + self.P, :]  # (planes_total?, P, 8)
        base_plane = (self._ptr - self.buf.data_ptr()) // 16 // self.ps
        body = self.buf[base_plane:base_plane + self.planes, self.lead:self.lead + self.P, :]
        t = body.reshape(self.planes, self.N, self.Hp, self.Wp, 8)[:, :, 1:, 1:, :]
        return t.permute(1, 0, 4, 2, 3).reshape(self.N, self.C, self.H, self.W).float()

    def padding_is_zero(self):
        base_plane = (self._ptr - self.buf.data_ptr()) // 16 // self.ps
        pl = self.buf[base_plane:base_plane + self.planes]
        body = pl[:, self.lead:self.lead + self.P, :].reshape(self.planes, self.N, self.Hp, self.Wp, 8)
        ok = bool((body[:, :, 0, :, :] == 0).all()) and bool((body[:, :, :, 0, :] == 0).all())
        ok = ok and bool((pl[:, :self.lead] == 0).all()) and bool((pl[:, self.lead + self.P:] == 0).all())
        return ok


class PhasePF8:
    """A [N, C, H, W] activation stored as 4 half-resolution PF8 tensors (phase (a,b) = X[2y+a, 2x+b], index 2a+b):
    the input format of the 3x3 stride-2 convs on the flat-shift path (include/hrnb.h, HRNB_CONV_IN_PHASES)."""
    __slots__ = ("buf", "N", "C", "H", "W", "half", "phase_stride", "_ptr")

    def __init__(self, N, C_, H, W, device="cuda"):
        assert C_ % 8 == 0 and H % 2 == 0 and W % 2 == 0
        self.N, self.C, self.H, self.W = N, C_, H, W
        g = PF8.__new__(PF8)               # geometry helper of one phase (no storage of its own)
        g.N, g.C, g.H, g.W = N, C_, H // 2, W // 2
        g.Hp, g.Wp = g.H + 1, g.W + 1
        g.P = N * g.Hp * g.Wp
        g.planes = C_ // 8
        g.lead = _lib.guard_lead(g.Wp)
        g.ps = g.lead + (g.P + 7) // 8 * 8 + _lib.guard_tail(g.Wp)
        self.half = g
        self.buf = _alloc((4, g.planes, g.ps, 8), device)
        self.phase_stride = g.planes * g.ps * 8
        self._ptr = self.buf.data_ptr() + g.lead * 16

    @property
    def ptr(self):
        return self._ptr

    @property
    def ps(self):
        return self.half.ps

    def to_nchw(self):
        """[N, C, H, W] float32 reassembled with torch indexing (test helper)."""
        g = self.half
        body = self.buf[:, :, g.lead:g.lead + g.P, :].reshape(4, g.planes, self.N, g.Hp, g.Wp, 8)[:, :, :, 1:, 1:, :]
        t = body.permute(2, 1, 5, 3, 4, 0).reshape(self.N, self.C, g.H, g.W, 2, 2)      # n c y x a b
        return t.permute(0, 1, 2, 4, 3, 5).reshape(self.N, self.C, self.H, self.W).float()

    def padding_is_zero(self):
        g = self.half
        body = self.buf[:, :, g.lead:g.lead + g.P, :].reshape(4, g.planes, self.N, g.Hp, g.Wp, 8)
        ok = bool((body[:, :, :, 0] == 0).all()) and bool((body[:, :, :, :, 0] == 0).all())
        return ok and bool((self.buf[:, :, :g.lead] == 0).all()) and bool((self.buf[:, :, g.lead + g.P:] == 0).all())
