"""Quantised embedding tables: ``Embedding8bit``, ``Embedding4bit`` (+ ``EmbeddingNF4`` / ``EmbeddingFP4``) and the
32-bit-optimizer-state convenience classes ``StableEmbedding`` / ``Embedding``.

Contract of the reference (bitsandbytes/nn/modules.py: StableEmbedding :28-131, Embedding :134-210,
Embedding8bit :833-877, Embedding4bit :880-977, EmbeddingFP4 / EmbeddingNF4 :980-1015): the table is quantised when
the module is moved to the device, exactly like Linear8bitLt / Linear4bit (same Int8Params / Params4bit
parameters), and a look-up dequantises ONLY the rows it gathers -- int8 rows scaled by their row statistic / 127,
or the gathered packed rows + their per-block absmax through ``dequantize_4bit`` (the blockwise kernel of the
hot path).  Saving these modules is not supported by the reference either.
"""
from __future__ import annotations

import copy
import logging
from typing import Optional

import torch
from torch import Tensor, nn

from .. import functional as F
from .modules import Int8Params, Params4bit, fix_4bit_weight_quant_state_from_module

logger = logging.getLogger(__name__)


class StableEmbedding(nn.Embedding):
    """Embedding + LayerNorm with Xavier-uniform initialisation (reference :28-131).  The reference also pins the
    optimizer state of this weight to 32 bits through its GlobalOptimManager; the 8-bit optimizers are outside this
    package's scope (f-4), so there is nothing to pin here."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 max_norm: Optional[float] = None, norm_type: float = 2.0, scale_grad_by_freq: bool = False,
                 sparse: bool = False, _weight: Optional[Tensor] = None, device=None, dtype=None) -> None:
        super().__init__(num_embeddings, embedding_dim, padding_idx, max_norm, norm_type, scale_grad_by_freq, sparse,
                         _weight, device, dtype)
        self.norm = nn.LayerNorm(embedding_dim, device=device)

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.weight)
        self._fill_padding_idx_with_zero()

    def _fill_padding_idx_with_zero(self) -> None:
        if self.padding_idx is not None:
            with torch.no_grad():
                self.weight[self.padding_idx].fill_(0)

    def forward(self, input: Tensor) -> Tensor:
        emb = nn.functional.embedding(input, self.weight, self.padding_idx, self.max_norm, self.norm_type,
                                      self.scale_grad_by_freq, self.sparse)
        # always normalise in fp32, hand back the table's dtype
        emb = emb.to(torch.get_default_dtype())
        return self.norm(emb).to(self.weight.dtype)


class Embedding(nn.Embedding):
    """nn.Embedding with Xavier-uniform initialisation (reference :134-210; see StableEmbedding for the optimizer note)."""

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.weight)
        if self.padding_idx is not None:
            with torch.no_grad():
                self.weight[self.padding_idx].fill_(0)


class Embedding8bit(nn.Embedding):
    """LLM.int8() row-quantised embedding table: int8 rows + one fp32 statistic per row (reference :833-877)."""

    def __init__(self, num_embeddings, embedding_dim, device=None, dtype=None):
        super().__init__(num_embeddings, embedding_dim, device=device, dtype=dtype)
        self.dtype = self.weight.data.dtype
        self.weight = Int8Params(self.weight.data, has_fp16_weights=False, requires_grad=False)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        raise NotImplementedError("Saving Embedding8bit module is not implemented")

    def forward(self, input: Tensor) -> Tensor:
        if getattr(self.weight, "SCB", None) is None:
            raise RuntimeError("Embedding layer is not quantized. Please call .cuda() or .to(device) first.")
        table, stats = self.weight.data, self.weight.SCB
        if table.shape != (self.num_embeddings, self.embedding_dim) or stats.shape != (self.num_embeddings,):
            raise RuntimeError("Embedding8bit: the quantised table does not match the module's shape")
        rows = nn.functional.embedding(input, table)
        scale = nn.functional.embedding(input, stats.view(self.num_embeddings, 1))
        return (rows * (scale / 127.0)).to(self.dtype)


class Embedding4bit(nn.Embedding):
    """NF4 / FP4 blockwise-quantised embedding table (reference :880-977)."""

    def __init__(self, num_embeddings, embedding_dim, dtype=None, quant_type="fp4", quant_storage=torch.uint8, device=None):
        super().__init__(num_embeddings, embedding_dim, device=device, dtype=dtype)
        self.dtype = self.weight.data.dtype
        self.weight = Params4bit(self.weight.data, requires_grad=False, compress_statistics=None, quant_type=quant_type,
                                 quant_storage=quant_storage, module=self)
        if embedding_dim % self.weight.blocksize != 0:
            logger.warning("Embedding size %d is not divisible by block size %d. This will lead to slow inference.",
                           embedding_dim, self.weight.blocksize)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        raise NotImplementedError("Saving Embedding4bit module is not implemented")

    def _gather_then_dequantize(self, input: Tensor) -> Tensor:
        """Rows are whole numbers of quantisation blocks: gather the packed bytes and the absmax of the looked-up
        rows, then dequantise just those (one blockwise-kernel launch over input.numel() * embedding_dim values)."""
        qs = self.weight.quant_state
        half = self.embedding_dim // 2
        per_row = self.embedding_dim // qs.blocksize
        packed = self.weight.data.view(torch.uint8).view(self.num_embeddings, half)
        picked = nn.functional.embedding(input, packed).reshape(-1, 1)
        absmax = nn.functional.embedding(input, qs.absmax.view(self.num_embeddings, per_row)).reshape(-1)
        sub = copy.deepcopy(qs)
        sub.absmax = absmax
        sub.shape = torch.Size((*input.shape, self.embedding_dim))
        return F.dequantize_4bit(picked, sub).to(self.dtype)

    def forward(self, input: Tensor) -> Tensor:
        fix_4bit_weight_quant_state_from_module(self)
        if self.embedding_dim % self.weight.quant_state.blocksize == 0:
            return self._gather_then_dequantize(input)
        table = F.dequantize_4bit(self.weight.data, self.weight.quant_state)
        return nn.functional.embedding(input, table).to(self.dtype)


class EmbeddingFP4(Embedding4bit):
    def __init__(self, num_embeddings, embedding_dim, dtype=None, quant_storage=torch.uint8, device=None):
        super().__init__(num_embeddings, embedding_dim, dtype=dtype, quant_type="fp4", quant_storage=quant_storage,
                         device=device)


class EmbeddingNF4(Embedding4bit):
    def __init__(self, num_embeddings, embedding_dim, dtype=None, quant_storage=torch.uint8, device=None):
        super().__init__(num_embeddings, embedding_dim, dtype=dtype, quant_type="nf4", quant_storage=quant_storage,
                         device=device)
