"""Multi-GPU partitioning of the STFT path.  One process per GPU,
torch.distributed for the plumbing (NCCL over NVLink on the B200 box, gloo in the CPU tests).

* Batched independent signals: `shard_batch` -- contiguous ranges of the batch per rank, no
  communication at all.
* One long stream: `stream_stft` / `stream_istft` -- rank d owns frames [F*d/G, F*(d+1)/G) and
  the samples [f0*hop, f1*hop) (the last rank up to n).  This is the multi-process flavour of the C
  library's vv_dsp_stft_stream_* handle (include/vv_dsp/b200.h) and uses the same scheme: the only
  communication is two sample halos of nfft-hop floats per boundary (12 KB at nfft=4096/hop=1024), one
  from each neighbour; the LEFT halo makes a rank also compute the nfft/hop-1 frames in front of its
  range, which the synthesis kernel (vv_dsp_stft_shard_inverse) re-synthesises for their overlap into
  the rank's first samples.  No partial sums are exchanged and the ranks' outputs concatenate to the
  bit-identical result of the unsharded call.  (Sizes without a marching kernel -- hop not dividing
  nfft, nfft outside 512..8192 -- fall back to the round-1 scheme: right halo only, trailing partial
  sums sent to the right neighbour, which adds and renormalises; equal to rounding, not bit for bit.)

The functions take a `vv_dsp_b200.Stft` handle; tensors may be CUDA tensors (device-resident
path) or CPU tensors (host-staged path).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def shard_batch(batch: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous [start, stop) of the batch dimension owned by `rank`"""
    return batch * rank // world, batch * (rank + 1) // world


def frame_range(frames: int, world: int, rank: int) -> tuple[int, int]:
    return frames * rank // world, frames * (rank + 1) // world


def owned_samples(n: int, nfft: int, hop: int, world: int, rank: int) -> tuple[int, int]:
    """[s0, s1) of the stream that `rank` owns: its frames' hop-blocks; the last rank also the rest"""
    frames = 0 if n < nfft else 1 + (n - nfft) // hop
    f0, f1 = frame_range(frames, world, rank)
    return f0 * hop, (n if rank == world - 1 else f1 * hop)


def window_sum(w2: torch.Tensor, nfft: int, hop: int, frames: int, t0: int, count: int) -> torch.Tensor:
    """sum over frames f in [0, frames) covering position t of w2[t - f*hop], t = t0 .. t0+count-1,
    accumulated in ascending frame order in float32 (the reference's norm_add order)."""
    t = torch.arange(t0, t0 + count, device=w2.device, dtype=torch.int64)
    acc = torch.zeros(count, device=w2.device, dtype=torch.float32)
    k = (nfft + hop - 1) // hop
    for i in range(k - 1, -1, -1):                      # descending i == ascending frame index
        f = torch.div(t, hop, rounding_mode="floor") - i
        p = t - f * hop
        ok = (f >= 0) & (f < frames) & (p < nfft)
        acc = acc + torch.where(ok, w2[p.clamp(0, nfft - 1)], torch.zeros((), device=w2.device))
    return acc


def _as_lib(x: torch.Tensor):
    return x if x.is_cuda else x.numpy()


def _from_lib(y, like: torch.Tensor) -> torch.Tensor:
    return y if isinstance(y, torch.Tensor) else torch.from_numpy(y)


def _exchange(send_to: int | None, send_buf: torch.Tensor | None, recv_from: int | None, recv_buf: torch.Tensor | None, group=None):
    ops = []
    if send_to is not None:
        ops.append(dist.P2POp(dist.isend, send_buf, send_to, group))
    if recv_from is not None:
        ops.append(dist.P2POp(dist.irecv, recv_buf, recv_from, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


def halo_mode(nfft: int, hop: int) -> bool:
    """True when (nfft, hop) has a marching synthesis kernel with shard mode (include/vv_dsp/b200.h)"""
    return nfft in (512, 1024, 2048, 4096, 8192) and hop in (nfft // 8, nfft // 4, nfft // 2)


def _is_emulator(h) -> bool:
    return "emu" in os.path.basename(getattr(h.lib, "path", ""))


def stream_stft_halo(h, x_owned: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """Analysis of this rank's shard, halo scheme: returns [halo_frames + own frames, bins] complex64 (the halo rows,
    nfft/hop - 1 of them on every rank but the first, duplicate the previous rank's last frames)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nfft, hop = h.nfft, h.hop
    frames = 0 if n < nfft else 1 + (n - nfft) // hop
    halo = nfft - hop
    assert frames // world >= nfft // hop, "every shard must hold at least nfft/hop frames"
    dev = x_owned.device
    left = torch.empty(halo, dtype=torch.float32, device=dev) if rank > 0 else None
    right = torch.empty(halo, dtype=torch.float32, device=dev) if rank < world - 1 else None
    if world > 1:
        ops = []
        if rank > 0:                     # my first samples are the left neighbour's right halo; its last ones my left halo
            ops += [dist.P2POp(dist.isend, x_owned[:halo].contiguous(), rank - 1, group), dist.P2POp(dist.irecv, left, rank - 1, group)]
        if rank < world - 1:
            ops += [dist.P2POp(dist.isend, x_owned[-halo:].contiguous(), rank + 1, group), dist.P2POp(dist.irecv, right, rank + 1, group)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    local = torch.cat([t for t in (left, x_owned, right) if t is not None])
    return _from_lib(h.batch_forward(_as_lib(local[None, :].contiguous()), "complex", "valid"), local)[0]


def stream_istft_halo(h, spec_local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """Synthesis of this rank's shard from stream_stft_halo's layout: the rank's owned samples, bit-identical to the
    same samples of the unsharded call.  No communication."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nfft, hop = h.nfft, h.hop
    s0, s1 = owned_samples(n, nfft, hop, world, rank)
    halo_frames = nfft // hop - 1 if rank > 0 else 0
    if spec_local.is_cuda:
        return h.shard_inverse(spec_local.contiguous(), halo_frames, rank == 0, rank == world - 1, s1 - s0)
    if _is_emulator(h):                                  # CPU tests: the emulator's "device" memory is host memory
        out = np.empty(s1 - s0, np.float32)
        h.shard_inverse_raw(np.ascontiguousarray(spec_local.numpy()), halo_frames, rank == 0, rank == world - 1, out)
        return torch.from_numpy(out)
    return h.shard_inverse(spec_local.cuda().contiguous(), halo_frames, rank == 0, rank == world - 1, s1 - s0).cpu()


def stream_stft(h, x_owned: torch.Tensor, n: int, kind: str = "complex", group=None) -> torch.Tensor:
    """STFT (valid frames) of one stream of n samples sharded by frame range.  `x_owned` is this
    rank's slice `owned_samples(...)`.  Returns this rank's frames [f1-f0, bins]."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nfft, hop = h.nfft, h.hop
    frames = 0 if n < nfft else 1 + (n - nfft) // hop
    f0, f1 = frame_range(frames, world, rank)
    halo = nfft - hop
    assert world == 1 or frames // world >= (nfft + hop - 1) // hop, "shards must be longer than one frame"
    recv = torch.empty(halo, dtype=torch.float32, device=x_owned.device) if rank < world - 1 else None
    send = x_owned[:halo].contiguous() if rank > 0 else None
    if halo and world > 1:
        _exchange(rank - 1 if rank > 0 else None, send, rank + 1 if rank < world - 1 else None, recv, group)
    local = torch.cat([x_owned, recv]) if (recv is not None and halo) else x_owned
    spec = _from_lib(h.batch_forward(_as_lib(local[None, :].contiguous()), kind, "valid"), local)[0]
    assert spec.shape[0] == f1 - f0, (spec.shape, f0, f1)
    return spec


def stream_istft(h, spec_local: torch.Tensor, n: int, window: torch.Tensor, group=None) -> torch.Tensor:
    """ISTFT with window-sum normalisation of a frame-range-sharded stream; `window` is the handle's
    window (float32[nfft], same device as spec_local).  Returns this rank's owned samples."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nfft, hop = h.nfft, h.hop
    frames = 0 if n < nfft else 1 + (n - nfft) // hop
    f0, f1 = frame_range(frames, world, rank)
    s0, s1 = owned_samples(n, nfft, hop, world, rank)
    fl, edge = f1 - f0, nfft - hop
    span = (fl - 1) * hop + nfft if fl else 0
    n_local = max(span, s1 - s0)
    y = _from_lib(h.batch_inverse(_as_lib(spec_local[None].contiguous()), n_local, True), spec_local)[0]
    if world > 1 and edge:
        w2 = (window * window).to(torch.float32)
        send = recv = None
        if rank < world - 1:                              # raw partial sums of the tail beyond my range
            send = (y[fl * hop: fl * hop + edge] * window_sum(w2, nfft, hop, fl, fl * hop, edge)).contiguous()
        if rank > 0:
            recv = torch.empty(edge, dtype=torch.float32, device=y.device)
        _exchange(rank + 1 if rank < world - 1 else None, send, rank - 1 if rank > 0 else None, recv, group)
        if rank > 0:                                      # head: undo the local normalisation, add, renormalise globally
            raw = y[:edge] * window_sum(w2, nfft, hop, fl, 0, edge) + recv
            norm = window_sum(w2, nfft, hop, frames, s0, edge)
            y[:edge] = torch.where(norm > 1e-12, raw / norm, torch.zeros_like(raw))
    return y[: s1 - s0].contiguous()


class StreamPlan:
    """Frame-range sharding of one stream with everything static hoisted out of the step: the local
    buffer [owned span | halo] is allocated once (the neighbour's halo is received straight into its
    tail, no concatenation), and the window-sum factors of the two nfft-hop edge regions are computed
    once.  Per step and rank: one halo recv/send, one fused STFT kernel, one fused ISTFT kernel, one
    tail send/recv and three small elementwise ops on nfft-hop samples.

        plan = StreamPlan(h, n, window)            # collective: every rank constructs it
        plan.x_owned[:] = my samples               # view into the local buffer
        spec = plan.stft()                         # [f1-f0, bins]
        y = plan.istft(spec)                       # [s1-s0]
    """

    def __init__(self, h, n: int, window: torch.Tensor, group=None):
        self.h, self.n, self.group = h, n, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nfft, hop = h.nfft, h.hop
        self.nfft, self.hop, self.edge = nfft, hop, nfft - hop
        self.frames = 0 if n < nfft else 1 + (n - nfft) // hop
        self.f0, self.f1 = frame_range(self.frames, self.world, self.rank)
        self.s0, self.s1 = owned_samples(n, nfft, hop, self.world, self.rank)
        self.fl = self.f1 - self.f0
        assert self.world == 1 or self.frames // self.world >= (nfft + hop - 1) // hop, "shards must be longer than one frame"
        dev = window.device
        own = self.s1 - self.s0
        self.has_right = self.rank < self.world - 1 and self.edge > 0
        self.has_left = self.rank > 0 and self.edge > 0
        self.local = torch.zeros(own + (self.edge if self.has_right else 0), dtype=torch.float32, device=dev)
        self.x_owned = self.local[:own]
        self.halo = self.local[own:] if self.has_right else None
        span = (self.fl - 1) * hop + nfft if self.fl else 0
        self.n_local = max(span, own)
        self.y = torch.empty((1, self.n_local), dtype=torch.float32, device=dev)
        self.bins = nfft // 2 + 1
        self.spec = torch.empty((1, self.fl, self.bins), dtype=torch.complex64, device=dev)
        w2 = (window * window).to(torch.float32)
        if self.has_right:       # raw tail = y_tail * local window-sum there
            self.tail_norm = window_sum(w2, nfft, hop, self.fl, self.fl * hop, self.edge)
            self.tail_send = torch.empty(self.edge, dtype=torch.float32, device=dev)
        if self.has_left:
            self.head_norm = window_sum(w2, nfft, hop, self.fl, 0, self.edge)
            gn = window_sum(w2, nfft, hop, self.frames, self.s0, self.edge)
            self.head_inv = torch.where(gn > 1e-12, 1.0 / gn, torch.zeros_like(gn))
            self.tail_recv = torch.empty(self.edge, dtype=torch.float32, device=dev)
            self.head_send = torch.empty(self.edge, dtype=torch.float32, device=dev)

    def stft(self, kind: str = "complex") -> torch.Tensor:
        if self.world > 1 and self.edge:
            if self.has_left:
                self.head_send.copy_(self.x_owned[: self.edge])
            _exchange(self.rank - 1 if self.has_left else None, self.head_send if self.has_left else None,
                      self.rank + 1 if self.has_right else None, self.halo, self.group)
        if kind == "complex":
            self.h.batch_forward(self.local[None, :], "complex", "valid", out=self.spec)
            return self.spec[0]
        return self.h.batch_forward(self.local[None, :], kind, "valid")[0]

    def istft(self, spec_local: torch.Tensor) -> torch.Tensor:
        self.h.batch_inverse(spec_local[None], self.n_local, True, out=self.y)
        y = self.y[0]
        if self.world > 1 and self.edge:
            if self.has_right:
                torch.mul(y[self.fl * self.hop: self.fl * self.hop + self.edge], self.tail_norm, out=self.tail_send)
            _exchange(self.rank + 1 if self.has_right else None, self.tail_send if self.has_right else None,
                      self.rank - 1 if self.has_left else None, self.tail_recv if self.has_left else None, self.group)
            if self.has_left:
                head = y[: self.edge]
                head.mul_(self.head_norm).add_(self.tail_recv).mul_(self.head_inv)
        return y[: self.s1 - self.s0]
