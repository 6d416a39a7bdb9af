"""4-bit quantisation of arbitrary module parameters through ``torch.nn.utils.parametrize`` (MoE expert tensors
and other weights that are not nn.Linear).

Contract of the reference's bitsandbytes/nn/parametrize.py (Bnb4bitParametrization :11-39,
replace_parameter_4bit_prequantized :42-59, replace_parameter_4bit :62-126, cache hooks :129-167, state-dict hook
:170-206): the parameter is replaced by its packed 4-bit bytes, a parametrization dequantises it on access
(``F.dequantize_4bit`` -- the blockwise kernel of the hot path), the dequantised tensor is cached for the duration
of the owning module's forward, and ``state_dict()`` stores the packed bytes under the plain parameter name with
the QuantState's packed entries next to it.
"""
from __future__ import annotations

from functools import partial
from typing import Any, Literal, Optional

import torch
import torch.nn as nn
import torch.nn.utils.parametrize as P

from .. import functional as F


class Bnb4bitParametrization(nn.Module):
    """Dequantises the (already packed) parameter whenever it is accessed."""

    def __init__(self, quant_state: F.QuantState):
        super().__init__()
        self.quant_state = quant_state

    @torch.no_grad()
    def forward(self, quantized_param: torch.Tensor) -> torch.Tensor:
        return F.dequantize_4bit(quantized_param, self.quant_state)


def _checked_parameter(module: nn.Module, param_name: str) -> nn.Parameter:
    if not hasattr(module, param_name):
        raise AttributeError(f"Module does not have parameter '{param_name}'")
    param = getattr(module, param_name)
    if not isinstance(param, nn.Parameter):
        raise TypeError(f"Parameter '{param_name}' is not an instance of nn.Parameter")
    return param


def _attach(module: nn.Module, param_name: str, quant_state: F.QuantState) -> None:
    # unsafe=True: the parametrization changes shape and dtype (packed uint8 -> the original tensor)
    P.register_parametrization(module, param_name, Bnb4bitParametrization(quant_state), unsafe=True)
    _register_parametrization_hooks(module, param_name)


def replace_parameter_4bit_prequantized(module: nn.Module, param_name: str, qs_dict: dict[str, Any],
                                        device: torch.device) -> None:
    """The parameter already holds packed bytes (a loaded checkpoint): attach the parametrization only."""
    _checked_parameter(module, param_name)
    _attach(module, param_name, F.QuantState.from_dict(qs_dict, device=device))


def replace_parameter_4bit(module: nn.Module, param_name: str, compress_statistics: bool = False,
                           quant_type: Literal["nf4", "fp4"] = "nf4", blocksize: Optional[int] = None) -> None:
    """Quantise ``module.<param_name>`` to 4 bits in place and dequantise it transparently on access."""
    param = _checked_parameter(module, param_name)
    packed, quant_state = F.quantize_4bit(param.data, blocksize=blocksize, compress_statistics=compress_statistics,
                                          quant_type=quant_type)
    setattr(module, param_name, nn.Parameter(packed, requires_grad=False))
    del param
    _attach(module, param_name, quant_state)


# torch's parametrization cache is a process-global counter + dict: enable it around the owning module's forward so
# that a parameter read several times in one forward is dequantised once, and make sure the counter is released
# even when the forward aborts (activation checkpointing with use_reentrant=False stops the recompute by raising).
def _enable_parametrization_cache(module: nn.Module, inputs: tuple[Any, ...]):
    P._cache_enabled += 1


def _disable_parametrization_cache(module: nn.Module, inputs: tuple[Any, ...], output: Any):
    P._cache_enabled = max(0, P._cache_enabled - 1)  # never negative: a negative count would read as "enabled"
    if not P._cache_enabled:
        P._cache = {}


def _register_parametrization_hooks(module: nn.Module, param_name: str) -> None:
    module.register_state_dict_post_hook(partial(_parametrized_state_dict_post_hook, param_name=param_name))
    module.register_forward_pre_hook(_enable_parametrization_cache)
    module.register_forward_hook(_disable_parametrization_cache, always_call=True)


def _parametrized_state_dict_post_hook(module: nn.Module, state_dict: dict[str, Any], prefix: str, local_metadata: Any,
                                       *, param_name: str = "weight", **kwargs) -> None:
    """``parametrizations.<name>.original`` -> ``<name>`` plus the packed QuantState entries."""
    original_key = f"{prefix}parametrizations.{param_name}.original"
    if original_key not in state_dict:
        return
    state_dict[f"{prefix}{param_name}"] = state_dict.pop(original_key)
    found = [p for p in module.parametrizations[param_name] if isinstance(p, Bnb4bitParametrization)]
    if not found:
        raise RuntimeError(f"no 4-bit parametrization registered for '{param_name}'")
    quant_state = found[0].quant_state
    if quant_state is not None:
        for key, value in quant_state.as_dict(packed=True).items():
            state_dict[f"{prefix}{param_name}.{key}"] = value
