"""Column-sharded Linear4bit: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 /
NVSwitch) for the plumbing.  New functionality -- the reference is a single-device library
(SURVEY.md section 2.3); what it does provide is proof that byte-range sharding of a packed
4-bit weight is lossless (reference tests/test_linear4bit.py:256-283).

Sharding rule.  ``W[N, K]`` is packed row-major, quantisation blocks
run along K and never straddle rows when ``K % blocksize == 0``, so a row range
``[n0, n1)`` owns the contiguous byte range ``[n0*K/2, n1*K/2)`` and the contiguous
absmax range ``[n0*K/bs, n1*K/bs)``.  With double quantisation the 8-bit absmax codes are
sliced the same way and the level-2 statistics (one fp32 per 256 blocks), the level-2 code
book and the offset are addressed through the shard's first *global* block index, which must
be a multiple of 256 so that ``global_block >> 8`` stays aligned: rows per shard * K / bs %
256 == 0.  The weight is quantised ONCE globally and then sliced, never re-quantised per
shard, so every shard reproduces the single-GPU result bit for bit.

Forward: every rank holds the replicated activations ``x[M, K]``, computes its
``[M, N/world]`` slice with the fused kernel *directly into its columns of the full-width
output* (strided-output entry point of the C ABI) and the slices are exchanged with one
all-gather.  The gather is along the inner dimension of a row-major matrix, which
``all_gather_into_tensor`` cannot write in place, so the exchange runs on a ``[world, M,
N/world]`` staging buffer; ``gather_output=False`` hands back the local slice instead (what
a following row-parallel layer wants).

Fused exchange (``PeerGather`` + ``forward_fused``): the output lives in symmetric memory
(``torch.distributed._symmetric_memory``: one ``[M, N]`` buffer per rank, every rank holds the
peers' mappings) and the GEMM epilogue stores each output element into ITS columns of EVERY rank's
buffer -- the all-gather rides on the kernel's own stores over NVLink / NVSwitch, tile by tile,
there is no separate collective and no permute; one symmetric-memory barrier per step publishes
the result.  Buffers alternate between two slots so that a rank may start step i + 1 while a peer
still reads step i.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import functional as F
from .backends.cuda import gemm_4bit_into, gemm_4bit_multi_out


@dataclass
class Shard4bit:
    """The slice of a quantised [N, K] weight owned by one rank."""

    packed: torch.Tensor            # uint8 [rows*K/2]
    absmax: torch.Tensor            # fp32 [rows*K/bs]  (plain)  |  level-2 absmax slice (nested)
    absmax_8bit: Optional[torch.Tensor]
    absmax_code: Optional[torch.Tensor]
    absmax_offset: Optional[torch.Tensor]
    rows: int
    row0: int
    K: int
    blocksize: int
    quant_type: str


def shard_rows(N: int, world: int, rank: int) -> tuple[int, int]:
    if N % world != 0:
        raise ValueError(f"out_features ({N}) must be divisible by the world size ({world})")
    rows = N // world
    return rank * rows, rows


def slice_quantized_weight(packed: torch.Tensor, qs: F.QuantState, world: int, rank: int) -> Shard4bit:
    """Cut rank's row range out of a globally quantised weight (no re-quantisation)."""
    N, K = qs.shape
    bs = qs.blocksize
    if K % bs != 0:
        raise ValueError(f"in_features ({K}) must be a multiple of the blocksize ({bs}) to shard by rows")
    row0, rows = shard_rows(N, world, rank)
    flat = packed.reshape(-1).view(torch.uint8) if packed.dtype != torch.uint8 else packed.reshape(-1)
    b0, b1 = row0 * K // 2, (row0 + rows) * K // 2
    a0, a1 = row0 * K // bs, (row0 + rows) * K // bs
    if qs.nested:
        if a0 % 256 != 0 or (a1 - a0) % 256 != 0:
            raise ValueError("double-quantised shards must start and end on a 256-block boundary "
                             f"(rows*K/blocksize = {a1 - a0})")
        return Shard4bit(packed=flat[b0:b1].contiguous(), absmax=qs.state2.absmax[a0 // 256:a1 // 256].contiguous(),
                         absmax_8bit=qs.absmax[a0:a1].contiguous(), absmax_code=qs.state2.code,
                         absmax_offset=qs.offset.reshape(1).float(), rows=rows, row0=row0, K=K, blocksize=bs,
                         quant_type=qs.quant_type)
    return Shard4bit(packed=flat[b0:b1].contiguous(), absmax=qs.absmax[a0:a1].contiguous(), absmax_8bit=None,
                     absmax_code=None, absmax_offset=None, rows=rows, row0=row0, K=K, blocksize=bs,
                     quant_type=qs.quant_type)


class ColumnParallelLinear4bit(torch.nn.Module):
    """``y = x @ dequant(W)^T + b`` with W's output features split across the process group."""

    def __init__(self, shard: Shard4bit, out_features: int, bias: Optional[torch.Tensor] = None,
                 group: Optional[dist.ProcessGroup] = None, gather_output: bool = True):
        super().__init__()
        self.shard = shard
        self.out_features = out_features
        self.group = group
        self.gather_output = gather_output
        self.bias_shard = None if bias is None else bias[shard.row0:shard.row0 + shard.rows].contiguous()
        self._stage = None

    @classmethod
    def from_quantized(cls, packed, qs: F.QuantState, bias=None, group=None, gather_output=True):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        return cls(slice_quantized_weight(packed, qs, world, rank), qs.shape[0], bias, group, gather_output)

    def local_forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, ldc: Optional[int] = None):
        """This rank's [M, rows] slice; written into ``out`` (row stride ``ldc`` elements) if given."""
        s = self.shard
        M = x.numel() // s.K
        if out is None:
            out = torch.empty((M, s.rows), device=x.device, dtype=x.dtype)
            ldc = s.rows
        gemm_4bit_into(x, s.packed, (s.rows, s.K), s.absmax, s.blocksize, s.quant_type, self.bias_shard, s.absmax_8bit,
                       s.absmax_code, s.absmax_offset, out, ldc)
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        s = self.shard
        lead = x.shape[:-1]
        M = x.numel() // s.K
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1 or not self.gather_output:
            return self.local_forward(x).view(*lead, s.rows)
        if self._stage is None or self._stage.shape[1] != M or self._stage.dtype != x.dtype:
            self._stage = torch.empty((world, M, s.rows), device=x.device, dtype=x.dtype)
        rank = dist.get_rank(self.group)
        self.local_forward(x, self._stage[rank], s.rows)
        dist.all_gather_into_tensor(self._stage.view(-1), self._stage[rank].reshape(-1), group=self.group)
        # [world, M, rows] -> [M, world*rows]
        return self._stage.permute(1, 0, 2).reshape(*lead, world * s.rows)


class PeerGather:
    """Two symmetric-memory ``[M, N]`` output slots shared by the ranks of ``group``."""

    def __init__(self, M: int, N: int, dtype: torch.dtype, device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.M, self.N, self.dtype = M, N, dtype
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty((M, N), dtype=dtype, device=device)
            self.handles.append(symm_mem.rendezvous(t, group))
            self.bufs.append(t)
        self.world = self.handles[0].world_size
        self.rank = self.handles[0].rank
        self.step = 0

    def slot(self):
        """(local tensor, [base address of that slot on rank r for every r], handle) of the next step."""
        i = self.step & 1
        self.step += 1
        return self.bufs[i], [int(p) for p in self.handles[i].buffer_ptrs], self.handles[i]


def fused_forward(layer: "ColumnParallelLinear4bit", x: torch.Tensor, peers: PeerGather) -> torch.Tensor:
    """``layer(x)`` with the all-gather fused into the GEMM epilogue; returns this rank's [M, N] slot."""
    s = layer.shard
    M = x.numel() // s.K
    if M != peers.M or layer.out_features != peers.N or x.dtype != peers.dtype:
        raise ValueError("PeerGather was built for a different output shape / dtype")
    local, bases, handle = peers.slot()
    col_bytes = s.row0 * local.element_size()
    # own buffer first, then the peers
    order = [peers.rank] + [r for r in range(peers.world) if r != peers.rank]
    ptrs = [bases[r] + col_bytes for r in order]
    ok = gemm_4bit_multi_out(x, s.packed, (s.rows, s.K), s.absmax, s.blocksize, s.quant_type, layer.bias_shard,
                             s.absmax_8bit, s.absmax_code, s.absmax_offset, ptrs, peers.N)
    if not ok:  # shape outside the tensor-core kernel: local slice + NCCL all-gather into the same slot
        stage = torch.empty((peers.world, M, s.rows), device=x.device, dtype=x.dtype)
        layer.local_forward(x, stage[peers.rank], s.rows)
        dist.all_gather_into_tensor(stage.view(-1), stage[peers.rank].reshape(-1), group=layer.group)
        local.copy_(stage.permute(1, 0, 2).reshape(M, peers.N))
    handle.barrier(channel=0)  # every rank's stores have landed everywhere
    return local


def reassemble_shards(shards: list[Shard4bit]) -> tuple[torch.Tensor, torch.Tensor]:
    """Inverse of slice_quantized_weight for the plain (non-nested) case: (packed, absmax)."""
    return torch.cat([s.packed for s in shards]), torch.cat([s.absmax for s in shards])
