"""The lift as dispatcher-visible operators: ``torch.ops.fiery_b200.lift_splat`` / ``lift_splat_backward``.

``torch.library.custom_op`` registrations on top of the same C ABI (libfiery_b200.so), so the fused lift is an operator the
dispatcher knows: it has a fake (meta) implementation for tracing / ``torch.compile``, an autograd formula registered with
``register_autograd`` (no Python ``autograd.Function`` in the graph), and an autocast rule that mirrors the reference -- under AMP
the reference's softmax and outer product run in fp32 (fiery/models/encoder.py:99-100 under autocast), so the operator's inputs are
cast to fp32.

The operators take plain tensors plus an integer ``handle`` naming the ``LiftSplat`` module that holds the frustum / BEV-grid
constants (a registry of weak references; the constants are tiny host-derived tensors, not operator inputs).
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch

_REGISTRY = {}          # handle -> (weakref to the LiftSplat module, (C, X, Y, channels_last) as python values for the fake impl)


def register_module(module, device: torch.device) -> int:
    handle = id(module)
    X, Y, _ = module._constants(device)["dim"]                      # cached host-side integers: no device sync here
    meta = (int(module.encoder_out_channels), int(X), int(Y), module.output_layout == "channels_last")
    entry = _REGISTRY.get(handle)
    if entry is None or entry[0]() is not module or entry[1] != meta:
        _REGISTRY[handle] = (weakref.ref(module, lambda _r, h=handle: _REGISTRY.pop(h, None)), meta)
    return handle


def _module(handle: int):
    entry = _REGISTRY.get(handle)
    m = entry[0]() if entry is not None else None
    if m is None:
        raise RuntimeError("fiery_b200::lift_splat: the LiftSplat module behind this handle is gone")
    return m


@torch.library.custom_op("fiery_b200::lift_splat", mutates_args=(), device_types="cuda")
def lift_splat(head: torch.Tensor, intrinsics: torch.Tensor, extrinsics: torch.Tensor, plan: Optional[torch.Tensor], handle: int,
               make_plan: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """(head (B'n, D+C, h, w), intrinsics (B', n, 3, 3), extrinsics (B', n, 4, 4)) -> (BEV (B', C, X, Y) fp32, plan).
    ``plan``: a geometry plan of this calibration, or None.  ``make_plan``: compute one (returned, for the backward) when none was
    passed; otherwise the second output is an empty tensor and the tile kernels evaluate the geometry themselves."""
    from . import lift as L
    m = _module(handle)
    native = head.dtype == torch.float32 or (head.dtype == torch.float16 and L.NATIVE_FP16_FORWARD)
    head_in = head if native else head.float()
    made = None
    if plan is None and make_plan and intrinsics.shape[0]:
        made = m.plan(intrinsics.to(head.device), extrinsics)
    out = m._launch_forward(head_in, intrinsics, extrinsics, plan=plan if plan is not None else made)
    # the second output is the plan MADE here (an operator output may not alias an input: a plan that was passed in is not returned)
    return out, (made if made is not None else torch.empty(0, dtype=torch.uint8, device=head.device))


@lift_splat.register_fake
def _(head, intrinsics, extrinsics, plan, handle, make_plan):
    C, X, Y, channels_last = _REGISTRY[handle][1]                   # python values only: nothing here touches a real tensor
    B = intrinsics.shape[0]
    bev = head.new_empty((B, X, Y, C), dtype=torch.float32).permute(0, 3, 1, 2) if channels_last \
        else head.new_empty((B, C, X, Y), dtype=torch.float32)
    return bev, head.new_empty((0,), dtype=torch.uint8)


@torch.library.custom_op("fiery_b200::lift_splat_backward", mutates_args=(), device_types="cuda")
def lift_splat_backward(head: torch.Tensor, intrinsics: torch.Tensor, extrinsics: torch.Tensor, grad_bev: torch.Tensor,
                        plan: Optional[torch.Tensor], handle: int) -> torch.Tensor:
    """Gradient of the BEV w.r.t. the head tensor (same shape and dtype as ``head``); the calibration gets none (geometry.py:300)."""
    m = _module(handle)
    h32 = head if head.dtype == torch.float32 else head.float()          # the backward kernel reads an fp32 head tensor
    g = m._launch_backward(h32, intrinsics, extrinsics, grad_bev, plan=plan if (plan is not None and plan.numel()) else None)
    return g if g.dtype == head.dtype else g.to(head.dtype)


@lift_splat_backward.register_fake
def _(head, intrinsics, extrinsics, grad_bev, plan, handle):
    return torch.empty_like(head)


def _setup_context(ctx, inputs, output):
    head, intrinsics, extrinsics, plan, handle, _make_plan = inputs
    _bev, plan_out = output
    ctx.handle = handle
    ctx.save_for_backward(head, intrinsics, extrinsics, plan if plan is not None else plan_out)


def _backward(ctx, grad_bev, _grad_plan):
    head, intrinsics, extrinsics, plan = ctx.saved_tensors
    grad_head = torch.ops.fiery_b200.lift_splat_backward(head, intrinsics, extrinsics, grad_bev, plan, ctx.handle)
    return grad_head, None, None, None, None, None


lift_splat.register_autograd(_backward, setup_context=_setup_context)
# AMP (baseline.yml PRECISION 16): the reference's softmax / outer product run in fp32 under autocast -> so does the operator
torch.library.register_autocast("fiery_b200::lift_splat", "cuda", torch.float32)
