"""Recurrent-head tape: ``Full_layer`` (GRU cell + output layer, models/rlmil.py:208-220) over all the calls of one
optimiser step, differentiated in ONE batched pass.

train_MuRCL.py:243,272 calls the same ``Full_layer`` object T x 2 times per optimiser step (both views of each of the T
patch-steps, one hidden-state chain through all of them) and backpropagates once at the end (:291-295).  Left to autograd,
the backward is a chain of ~150 launches on 128-row operands: every call's three weight gradients are separate small GEMMs
(each with its split-K reduction, column sum and accumulation), every operand is cast on its own.  The tape records the
forward calls into [calls, B, .] buffers and, given the loss gradients of all calls, runs

  * one input-gradient and one weight-gradient GEMM for the output layer over all calls (M = calls x B rows),
  * the unavoidable sequential part: per call one cell backward (fused with the sum of its incoming gradients) and one
    W_hh input-gradient GEMM,
  * one weight-gradient GEMM each for W_hh and W_ih and one input-gradient GEMM for W_ih over all calls.

The weight gradients go straight into the parameters' persistent gradient views (``ops.grad_target``) when they have one.
Numerically this is the same computation as the per-call autograd chain (same GEMM kernels, same cell formulas); only the
summation order over the calls differs.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Sequence

import torch

from . import _lib, ops
from ._lib import check


def _p(t):
    return None if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


class HeadTape:
    def __init__(self, fc, calls: int, n_views: int, batch: int, device, precision: str):
        if not getattr(fc, "fc_rnn", False):
            raise ops.MurclError("HeadTape implements Full_layer(fc_rnn=True)")
        self.fc, self.nv, self.B, self.n_calls = fc, int(n_views), int(batch), int(calls)
        self.H, self.F, self.C = fc.hidden_state_dim, fc.feature_num, fc.class_num
        self.dt = ops.storage_dtype(precision)
        self.code = _lib.BF16 if self.dt == torch.bfloat16 else _lib.F32
        n = self.n_calls * self.nv                       # chain steps
        B, H, F = self.B, self.H, self.F
        f32 = dict(device=device, dtype=torch.float32)
        st = dict(device=device, dtype=self.dt)
        self.Xs = torch.empty((n, B, F), **st)           # inputs in the GEMM storage type
        self.GI = torch.empty((n, B, 3 * H), **f32)
        self.GH = torch.empty((n, B, 3 * H), **f32)
        self.GATES = torch.empty((n, B, 3 * H), **f32)
        self.Hf = torch.empty((n, B, H), **f32)
        self.Hs = torch.empty((n, B, H), **st)           # h_new of every chain step (operand of the output layer)
        self.HPs = torch.zeros((n + 1, B, H), **st)      # h_prev of every chain step (zeros where the chain restarts)
        self.restart: List[bool] = []
        self.x_inputs: List[torch.Tensor] = []
        self.stacked_inputs: List[bool] = []                 # per call: one stacked input instead of one per view
        self.z_leaves: List[torch.Tensor] = []
        self.step = 0
        w_ih, w_hh, self.b_ih, self.b_hh = fc.rnn.weight_ih_l0, fc.rnn.weight_hh_l0, fc.rnn.bias_ih_l0, fc.rnn.bias_hh_l0
        self.params = (w_ih, w_hh, self.b_ih, self.b_hh, fc.fc.weight, fc.fc.bias)
        self.w_ih_s, self.w_hh_s, self.w_fc_s = (ops.weight_as(w, self.dt) for w in (w_ih, w_hh, fc.fc.weight))
        self.bias = [b.detach().contiguous().float() for b in (self.b_ih, self.b_hh, fc.fc.bias)]
        self.dec = None                                      # (weight, bias, w_storage, bias_f32) once attach_decoder() ran
        self.m_inputs: List[torch.Tensor] = []               # decoder inputs (pooled bag vectors), one per call

    # ---------------------------------------------------------------------------------------------------
    def attach_decoder(self, weight: torch.nn.Parameter, bias: torch.nn.Parameter) -> None:
        """Also tape the aggregator's ``decoder`` layer (``Linear + ReLU`` on the pooled bag vectors, abmil.py:31,44): the
        actor needs its OUTPUT at every patch-step, so the forward stays per call (``decode``); its backward - ReLU mask,
        input gradient, weight gradient, bias column sums - runs ONCE over all calls (M = calls x n_views x B rows) instead
        of once per patch-step at the head of every aggregator's backward."""
        if self.step != 0:
            raise ops.MurclError("HeadTape.attach_decoder: attach before the first call")
        if weight.shape[0] != self.F:
            raise ops.MurclError(f"HeadTape.attach_decoder: the decoder produces {weight.shape[0]} features, Full_layer takes {self.F}")
        dev = self.Xs.device
        rows = self.nv * self.B
        self.dec = (weight, bias, ops.weight_as(weight, self.dt), None if bias is None else bias.detach().contiguous().float())
        self.Ms = torch.empty((self.n_calls, rows, weight.shape[1]), device=dev, dtype=self.dt)       # decoder inputs, storage type
        self.DEC = torch.empty((self.n_calls, rows, self.F), device=dev, dtype=torch.float32)         # decoder outputs (post-ReLU)

    @torch.no_grad()
    def decode(self, M: torch.Tensor) -> torch.Tensor:
        """``relu(M W_dec^T + b_dec)`` of the CURRENT call (call it before ``forward_views``); ``M`` ``[n_views * B, L]`` is
        recorded as the call's differentiable input.  Returns the ``[n_views * B, F]`` fp32 output (no graph attached): hand
        it to ``forward_views(..., stacked=...)`` and, detached, to the actor."""
        if self.dec is None:
            raise ops.MurclError("HeadTape.decode: attach_decoder() first")
        if self.step >= self.n_calls or len(self.m_inputs) != self.step:
            raise ops.MurclError("HeadTape.decode: one decode() per call, before forward_views()")
        t = self.step
        src = M.detach().contiguous()
        if tuple(src.shape) != tuple(self.Ms[t].shape):
            raise ops.MurclError(f"HeadTape.decode: expected {tuple(self.Ms[t].shape)}, got {tuple(src.shape)}")
        if src.dtype == self.dt:
            self.Ms[t].copy_(src)
        else:
            ops.cast_into(src.float() if src.dtype != torch.float32 else src, self.Ms[t])
        ops.linear_fwd(self.Ms[t], self.dec[2], self.dec[3], ops.ACT_RELU, out=self.DEC[t])
        self.m_inputs.append(M)
        return self.DEC[t]

    @property
    def grad_roots(self) -> List[torch.Tensor]:
        """The tensors ``backward`` returns gradients for, in order: the decoder inputs when a decoder is taped, else the
        recorded ``forward_views`` inputs."""
        return self.m_inputs if self.dec is not None else self.x_inputs

    # ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_views(self, xs: Sequence[torch.Tensor], restart: bool = False, stacked: torch.Tensor = None) -> List[torch.Tensor]:
        """``[fc(x, restart) for x in xs]`` (train_MuRCL.py:243,272): returns one LEAF tensor per view (``requires_grad``);
        their gradients are what ``backward`` consumes.  ``stacked`` (``[n_views * B, F]``, the tensor the views ``xs`` are
        consecutive row blocks of): recorded as the call's ONE input instead of the views, so that ``backward`` returns a
        single gradient for it and autograd does not rebuild it from per-view slices (two fills and an add per call)."""
        nv, B, H = self.nv, self.B, self.H
        if len(xs) != nv or any(tuple(x.shape) != (B, self.F) for x in xs):
            raise ops.MurclError(f"HeadTape: expected {nv} views of shape {(B, self.F)}")
        if stacked is not None and tuple(stacked.shape) != (nv * B, self.F):
            raise ops.MurclError(f"HeadTape: stacked input must be {(nv * B, self.F)}")
        if self.step >= self.n_calls:
            raise ops.MurclError("HeadTape: more calls than the tape was sized for")
        if self.step == 0 and not restart:
            raise ops.MurclError("HeadTape: the first call of a tape must restart the hidden state (train_MuRCL.py:243)")
        lib = _lib.load()
        c0 = self.step * nv
        for v, x in enumerate([stacked] if stacked is not None else xs):
            src = x.detach().contiguous()
            dst = self.Xs[c0:c0 + nv] if stacked is not None else self.Xs[c0 + v]
            if src.dtype == self.dt:
                dst.view(src.shape).copy_(src)
            else:
                ops.cast_into(src.float() if src.dtype != torch.float32 else src, dst)
            self.x_inputs.append(x)
        self.stacked_inputs.append(stacked is not None)
        ops.linear_fwd(self.Xs[c0:c0 + nv].view(nv * B, self.F), self.w_ih_s, self.bias[0], out=self.GI[c0:c0 + nv].view(nv * B, 3 * H))
        if restart and self.step > 0:
            self.HPs[c0:c0 + nv + 1].zero_()                                   # (a fresh tape is zero already)
        for v in range(nv):
            c = c0 + v
            first = restart
            self.restart.append(first)
            ops.linear_fwd(self.HPs[c], self.w_hh_s, self.bias[1], out=self.GH[c])
            # the next call continues from this state unless the whole call restarts (every view then starts from zeros)
            nxt = None if (restart and v + 1 < nv) else self.HPs[c + 1]
            check(lib.murcl_gru_cell_fwd_tape(_p(self.GI[c]), _p(self.GH[c]), None if first else _p(self.Hf[c - 1]), _p(self.Hf[c]),
                                              _p(self.Hs[c]), _p(nxt), _p(self.GATES[c]), B, H, self.code, _s()),
                  "murcl_gru_cell_fwd_tape")
        z = ops.linear_fwd(self.Hs[c0:c0 + nv].view(nv * B, H), self.w_fc_s, self.bias[2], out_dtype=torch.float32)
        self.fc.hidden = self.Hf[c0 + nv - 1].unsqueeze(0)
        self.step += 1
        outs = []
        for v in range(nv):
            leaf = z[v * B:(v + 1) * B].detach().requires_grad_(True)
            self.z_leaves.append(leaf)
            outs.append(leaf)
        return outs

    # ---------------------------------------------------------------------------------------------------
    def _grad_into(self, param, want_shape):
        tgt = ops.grad_target(param)
        if tgt is not None:
            return tgt, False
        return torch.zeros(want_shape, device=param.device, dtype=torch.float32), True

    @torch.no_grad()
    def backward(self, dzs: Sequence[torch.Tensor], side=None) -> List[torch.Tensor]:
        """``dzs``: gradient of the loss w.r.t. every leaf returned by ``forward_views`` (same order).  Accumulates the
        gradients of the six Full_layer parameters and returns the gradient w.r.t. every recorded input (``x_inputs``: one per
        view, or one per call where ``stacked`` was given), to be pushed into the graph that produced them.  ``side`` (a CUDA stream): the three batched
        weight-gradient GEMMs - which nothing downstream of this call reads - are issued there, beside the recurrence and
        the start of the aggregators' backward; the CALLER joins it (``current_stream().wait_stream(side)``) before the
        parameter gradients are used.  The operands stay referenced by the tape until it is dropped."""
        nv, B, H, F, C = self.nv, self.B, self.H, self.F, self.C
        n = self.step * nv
        if len(dzs) != n:
            raise ops.MurclError(f"HeadTape.backward: {len(dzs)} gradients for {n} recorded views")
        lib = _lib.load()
        dev = self.Xs.device
        dz = torch.stack([g.float() for g in dzs], 0).contiguous()                   # [n, B, C]
        DZs = dz if self.dt == torch.float32 else ops.cast(dz.view(n * B, C), self.dt).view(n, B, C)
        w_ih, w_hh, b_ih, b_hh, w_fc, b_fc = self.params
        grads, own = {}, {}
        for name, prm in (("w_ih", w_ih), ("w_hh", w_hh), ("b_ih", b_ih), ("b_hh", b_hh), ("w_fc", w_fc), ("b_fc", b_fc)):
            grads[name], own[name] = self._grad_into(prm, prm.shape)
        # output layer over all calls
        if any(own.values()):
            side = None          # gradients returned through .grad are combined on the calling stream right below
        main = torch.cuda.current_stream()
        on_side = (lambda: torch.cuda.stream(side)) if side is not None else contextlib.nullcontext
        dHd = ops.linear_bwd_input(DZs.view(n * B, C), self.w_fc_s).view(n, B, H)
        if side is not None:
            side.wait_stream(main)
        with on_side():
            ops.linear_bwd_weight(DZs.view(n * B, C), self.Hs[:n].view(n * B, H), True, dw_into=grads["w_fc"], db_into=grads["b_fc"])
        # the recurrence, last call first
        DGI = torch.empty((n, B, 3 * H), device=dev, dtype=self.dt)
        DGH = torch.empty((n, B, 3 * H), device=dev, dtype=self.dt)
        carry_f = [torch.empty((B, H), device=dev, dtype=torch.float32) for _ in range(2)]
        carry_s = None
        have_carry = False
        accum = os.environ.get("MURCL_TAPE_ACCUM_DGRAD", "1") != "0"
        flip = 0
        for c in range(n - 1, -1, -1):
            first = self.restart[c]
            out_f = None if first else carry_f[flip ^ 1]
            check(lib.murcl_gru_cell_bwd_tape(_p(dHd[c]), _p(carry_s) if (have_carry and carry_s is not None) else None,
                                              _p(carry_f[flip]) if have_carry else None, _p(self.GATES[c]), _p(self.GH[c]),
                                              None if first else _p(self.Hf[c - 1]), _p(DGI[c]), _p(DGH[c]), _p(out_f), B, H,
                                              self.code, _s()), "murcl_gru_cell_bwd_tape")
            if first:
                have_carry = False
            else:
                # through gh = h_prev W_hh^T: added straight into the fp32 buffer that holds the cell's direct term (split-K
                # with atomics over the whole GPU) where the kernel takes the shape, else a storage-type tensor of its own
                carry_s = None if (accum and ops.linear_bwd_input_accum_(DGH[c], self.w_hh_s, out_f)) else \
                    ops.linear_bwd_input(DGH[c], self.w_hh_s)
                flip ^= 1
                have_carry = True
        # weight gradients of the recurrence over all calls (h_prev rows of restarting calls are zeros: exact)
        if side is not None:
            side.wait_stream(main)
            self._keep = (dz, DZs, DGI, DGH)            # read on the side stream: must outlive this call
        with on_side():
            ops.linear_bwd_weight(DGH.view(n * B, 3 * H), self.HPs[:n].view(n * B, H), True, dw_into=grads["w_hh"], db_into=grads["b_hh"])
            ops.linear_bwd_weight(DGI.view(n * B, 3 * H), self.Xs[:n].view(n * B, F), True, dw_into=grads["w_ih"], db_into=grads["b_ih"])
        dX = ops.linear_bwd_input(DGI.view(n * B, 3 * H), self.w_ih_s)
        dX = dX.float() if dX.dtype != torch.float32 else dX
        for name, prm in (("w_ih", w_ih), ("w_hh", w_hh), ("b_ih", b_ih), ("b_hh", b_hh), ("w_fc", w_fc), ("b_fc", b_fc)):
            if own[name]:
                prm.grad = grads[name] if prm.grad is None else prm.grad + grads[name]
        dX = dX.view(n, B, F)
        if self.dec is not None:
            return self._decoder_backward(dX, side, main)
        out = []
        for call, stacked in enumerate(self.stacked_inputs):     # same order as x_inputs
            if stacked:
                out.append(dX[call * nv:(call + 1) * nv].reshape(nv * B, F))
            else:
                out.extend(dX[call * nv + v] for v in range(nv))
        return out

    def _decoder_backward(self, dX: torch.Tensor, side, main) -> List[torch.Tensor]:
        """Backward of the taped decoder over ALL calls at once: ``dX`` ``[calls * n_views, B, F]`` (gradient w.r.t. the
        decoder outputs) -> one gradient per recorded decoder input."""
        if len(self.m_inputs) != self.step or not all(self.stacked_inputs):
            raise ops.MurclError("HeadTape: with a taped decoder every call is decode() + forward_views(stacked=its result)")
        weight, bias, w_s, _ = self.dec
        calls, rows = self.step, self.nv * self.B
        dpre = ops.relu_bwd(dX.reshape(calls * rows, self.F).contiguous(), self.DEC[:calls].view(calls * rows, self.F))
        dpre_s = dpre if self.dt == torch.float32 else ops.cast(dpre, self.dt)
        dM = ops.linear_bwd_input(dpre_s, w_s)
        dM = dM.float() if dM.dtype != torch.float32 else dM
        gw, own_w = self._grad_into(weight, weight.shape)
        gb, own_b = (None, False) if bias is None else self._grad_into(bias, bias.shape)
        if own_w or own_b:
            side = None
        if side is not None:
            side.wait_stream(main)
            self._keep_dec = (dpre_s,)                      # read on the side stream: must outlive this call
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            ops.linear_bwd_weight(dpre_s, self.Ms[:calls].view(calls * rows, -1), bias is not None, dw_into=gw, db_into=gb)
        if own_w:
            weight.grad = gw if weight.grad is None else weight.grad + gw
        if own_b:
            bias.grad = gb if bias.grad is None else bias.grad + gb
        dM = dM.view(calls, rows, -1)
        return [dM[t] for t in range(calls)]
