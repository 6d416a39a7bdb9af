"""Serialisation helpers for QuantState (reference bitsandbytes/utils.py:166-201).

The non-tensor fields of a QuantState travel inside a checkpoint as one uint8 tensor
holding UTF-8 JSON; this is the on-disk format of every ``bnb-4bit`` checkpoint on the
HF hub, so it must not change.
"""
from __future__ import annotations

import json
from typing import Any

import torch


def pack_dict_to_tensor(source_dict: dict[str, Any]) -> torch.Tensor:
    """JSON-encode ``source_dict`` into a 1-D uint8 tensor."""
    payload = json.dumps(source_dict).encode("utf-8")
    return torch.tensor(list(payload), dtype=torch.uint8)


def unpack_tensor_to_dict(tensor_data: torch.Tensor) -> dict[str, Any]:
    """Inverse of :func:`pack_dict_to_tensor`."""
    raw = bytes(tensor_data.detach().cpu().to(torch.uint8).tolist())
    return json.loads(raw.decode("utf-8"))


def sync_gpu(t):
    """Block until the device of `t` is idle (reference utils.py:204-208; CUDA only here)."""
    import torch

    if t.device.type == "cuda":
        torch.cuda.synchronize()
