"""Host side of the tcgen05 grouped GEMM (``fjsp_a2c_gemm``, csrc/fjsp_umma.cuh): problem tables.

A table is a device-resident array of ``FjspGemmProb`` records (include/fjsp_b200.h) built once over static buffers —
the trainer's rollout and update replay from CUDA graphs, so every pointer is fixed for the life of the trainer — and
launched with one call.  All problems of a table share the operand orientation pair.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi

OP_KC, OP_KCS, OP_MC, OP_PK = abi.OP_KC, abi.OP_KCS, abi.OP_MC, abi.OP_PK
RELU, ATOMIC = abi.GEMM_RELU, abi.GEMM_ATOMIC
BM = 128


def _addr(t, off=0):
    return 0 if t is None else t.data_ptr() + 4 * int(off)


class GemmTable:
    """C[M x N] (+)= A[M x K] @ B[N x K]^T for a list of problems, one launch.

    ``add`` takes tensors (fp32, on the table's device) plus element offsets into them, so slices of bigger buffers
    (an observation slice, one actor of a stacked parameter) need no copies.  Orientation ``a_op`` / ``b_op``:
    OP_KC / OP_KCS: X(r, k) = X[r * ld + k]; OP_MC: X(r, k) = X[k * ld + r]."""

    def __init__(self, device, a_op, b_op, passes=3):
        self.device, self.a_op, self.b_op, self.passes = torch.device(device), int(a_op), int(b_op), int(passes)
        self.rows, self.keep = [], []
        self.dev_table = None
        self.max_ctas = 1

    def add(self, A, B, Cm, M, N, K, lda, ldb, csm, csn=1, a_off=0, b_off=0, c_off=0, bias=None, bias_off=0, mask=None,
            mask_off=0, colsum=None, colsum_off=0, relu=False, atomic=False, splitk=1, rowdot_w=None, rowdot_out=None,
            rowdot_out_off=0, rowdot_bias=None, rowdot_w_off=0, rowdot_bias_off=0, head_n=0, head_ld=0):
        """rowdot_*: fused one-column head on the stored values: rowdot_out[m] = sum_n C[m, n] * rowdot_w[n] + rowdot_bias[0];
        with head_n = R in 1..8 a head of R columns (any N): rowdot_out[m * head_ld + j] = sum_n C[m, n] * rowdot_w[n * R + j] +
        rowdot_bias[j]."""
        assert 1 <= N <= 256 and M >= 1 and K >= 1 and splitk >= 1
        assert rowdot_w is None or (rowdot_out is not None and not atomic and (N <= 128 if head_n == 0 else 1 <= head_n <= 8))
        assert head_n == 0 or (rowdot_w is not None and head_ld >= head_n)
        for t in (A, B, Cm, bias, mask, colsum, rowdot_w, rowdot_out, rowdot_bias):
            assert t is None or (t.dtype == torch.float32 and t.device == self.device), "fp32 tensors on the table's device"
        if self.a_op == OP_KC:
            assert lda % 4 == 0 and K % 4 == 0 and _addr(A, a_off) % 16 == 0, "OP_KC needs 16-byte aligned rows"
        if self.b_op == OP_KC:
            assert ldb % 4 == 0 and K % 4 == 0 and _addr(B, b_off) % 16 == 0, "OP_KC needs 16-byte aligned rows"
        if self.b_op == OP_PK:   # B = a packed image (PackTable): ldb is not used
            assert _addr(B, b_off) % 16 == 0 and B.numel() - b_off >= abi.pack_image_floats(N, K), "packed image too small / unaligned"
        r = np.zeros((), dtype=abi.GEMM_PROB_DT)
        r["A"], r["B"], r["C"] = _addr(A, a_off), _addr(B, b_off), _addr(Cm, c_off)
        r["bias"], r["mask"], r["colsum"] = _addr(bias, bias_off), _addr(mask, mask_off), _addr(colsum, colsum_off)
        r["M"], r["N"], r["K"], r["lda"], r["ldb"], r["csm"], r["csn"] = M, N, K, lda, ldb, csm, csn
        r["flags"] = (RELU if relu else 0) | (ATOMIC if atomic else 0)
        r["splitk"] = splitk
        r["rowdot_w"], r["rowdot_out"], r["rowdot_bias"] = _addr(rowdot_w, rowdot_w_off), _addr(rowdot_out, rowdot_out_off), _addr(rowdot_bias, rowdot_bias_off)
        r["head_n"], r["head_ld"] = head_n, head_ld
        self.rows.append(r)
        self.keep += [A, B, Cm, bias, mask, colsum, rowdot_w, rowdot_out, rowdot_bias]
        self.max_ctas = max(self.max_ctas, (M + BM - 1) // BM * splitk)
        self.dev_table = None
        return self

    def finalize(self):
        host = np.stack(self.rows)
        self.dev_table = torch.from_numpy(host.view(np.uint8).reshape(len(self.rows), -1).copy()).to(self.device)
        return self

    def launch(self, stream=None, passes=None):
        if self.dev_table is None:
            self.finalize()
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        rc = abi.lib().fjsp_a2c_gemm(C.c_void_p(self.dev_table.data_ptr()), len(self.rows), self.max_ctas, self.a_op, self.b_op,
                                     self.passes if passes is None else int(passes), C.c_void_p(st))
        if rc:
            abi.check(rc)



class PackTable:
    """Weight matrices -> packed B images (``fjsp_a2c_gemm_pack``, include/fjsp_b200.h FJSP_OP_PK), one launch for all of
    them.  ``add`` returns (image tensor, element offset) to pass as ``B`` / ``b_off`` of a ``GemmTable`` with b_op = OP_PK."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.jobs, self.keep, self.total = [], [], 0
        self.image = None
        self.dev_table = None

    def add(self, W, op, ld, N, K, w_off=0):
        assert W.dtype == torch.float32 and W.device == self.device
        off = self.total
        self.total += abi.pack_image_floats(N, K)
        self.jobs.append((W, int(w_off), int(op), int(ld), int(N), int(K), off))
        return off

    def finalize(self):
        self.image = torch.zeros(max(self.total, 4), device=self.device)
        rows = np.zeros(len(self.jobs), dtype=abi.PACK_JOB_DT)
        for r, (W, w_off, op, ld, N, K, off) in zip(rows, self.jobs):
            r["src"], r["dst"], r["op"], r["ld"], r["N"], r["K"] = _addr(W, w_off), _addr(self.image, off), op, ld, N, K
        self.dev_table = torch.from_numpy(rows.view(np.uint8).reshape(len(self.jobs), -1).copy()).to(self.device)
        return self

    def launch(self, stream=None):
        """(Re)build every image from the current values of the weights."""
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        rc = abi.lib().fjsp_a2c_gemm_pack(C.c_void_p(self.dev_table.data_ptr()), len(self.jobs), C.c_void_p(st))
        if rc:
            abi.check(rc)


class WgradTable:
    """Narrow weight gradients (``fjsp_a2c_wgrad_small``): G[i, j] += sum_b X[b, i] * Y[b, j] with nx <= 256, ny <= 40, plain
    fp32.  Jobs are grouped by the width of Y (<= 8, <= 16, <= 40) into at most three launches."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.jobs, self.keep, self.groups = [], [], None

    def add(self, X, Y, G, B, nx, ny, ldx, ldy, gsi, gsj, x_off=0, y_off=0, g_off=0):
        assert 1 <= nx <= 256 and 1 <= ny <= 40 and B >= 1
        for t in (X, Y, G):
            assert t.dtype == torch.float32 and t.device == self.device
        r = np.zeros((), dtype=abi.WGRAD_JOB_DT)
        r["X"], r["Y"], r["G"] = _addr(X, x_off), _addr(Y, y_off), _addr(G, g_off)
        r["B"], r["nx"], r["ny"], r["ldx"], r["ldy"], r["gsi"], r["gsj"] = B, nx, ny, ldx, ldy, gsi, gsj
        self.jobs.append(r)
        self.keep += [X, Y, G]
        self.groups = None
        return self

    def finalize(self):
        self.groups = []
        for lo, hi in ((1, 8), (9, 16), (17, 40)):
            rows = [r for r in self.jobs if lo <= int(r["ny"]) <= hi]
            if rows:
                host = np.stack(rows)
                tab = torch.from_numpy(host.view(np.uint8).reshape(len(rows), -1).copy()).to(self.device)
                self.groups.append((tab, len(rows), max(int(r["B"]) for r in rows), hi))
        return self

    def launch(self, stream=None):
        if self.groups is None:
            self.finalize()
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        for tab, n, max_rows, max_ny in self.groups:
            rc = abi.lib().fjsp_a2c_wgrad_small(C.c_void_p(tab.data_ptr()), n, max_rows, max_ny, C.c_void_p(st))
            if rc:
                abi.check(rc)


class Layer1Table:
    """First layers with a short K (``fjsp_a2c_layer1``): Y = relu?(X W + b), K <= 40, N <= 256, plain fp32 FMAs, one job per
    network, one launch."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.rows, self.keep, self.dev_table, self.max_rows, self.max_k = [], [], None, 1, 1

    def add(self, X, W, Y, rows, n, k, ldx, ldy, x_off=0, w_off=0, y_off=0, bias=None, bias_off=0, relu=True):
        assert 1 <= k <= 40 and 1 <= n <= 256 and rows >= 1
        for t in (X, W, Y, bias):
            assert t is None or (t.dtype == torch.float32 and t.device == self.device)
        r = np.zeros((), dtype=abi.LAYER1_JOB_DT)
        r["X"], r["W"], r["bias"], r["Y"] = _addr(X, x_off), _addr(W, w_off), _addr(bias, bias_off), _addr(Y, y_off)
        r["rows"], r["k"], r["n"], r["ldx"], r["ldy"], r["relu"] = rows, k, n, ldx, ldy, int(bool(relu))
        self.rows.append(r)
        self.keep += [X, W, Y, bias]
        self.max_rows, self.max_k = max(self.max_rows, int(rows)), max(self.max_k, int(k))
        self.dev_table = None
        return self

    def finalize(self):
        host = np.stack(self.rows)
        self.dev_table = torch.from_numpy(host.view(np.uint8).reshape(len(self.rows), -1).copy()).to(self.device)
        return self

    def launch(self, stream=None):
        if self.dev_table is None:
            self.finalize()
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        rc = abi.lib().fjsp_a2c_layer1(C.c_void_p(self.dev_table.data_ptr()), len(self.rows), self.max_rows, self.max_k, C.c_void_p(st))
        if rc:
            abi.check(rc)


class HeadBwdTable:
    """Backward through narrow heads (``fjsp_a2c_head_backward``): per job dH = (dl W^T) * (H > 0), gb += column sums of dH,
    gW += H^T dl, in one pass over H; one launch for all jobs."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.rows, self.keep, self.dev_table, self.max_rows = [], [], None, 1

    def add(self, dl, W, H, dH, gW, gb, rows, n, na, ld_dl, dl_off=0, w_off=0, h_off=0, dh_off=0, gw_off=0, gb_off=0):
        assert 1 <= na <= 8 and 4 <= n <= 256 and n % 4 == 0 and rows >= 1
        for t in (dl, W, H, dH, gW, gb):
            assert t is None or (t.dtype == torch.float32 and t.device == self.device)
        assert _addr(H, h_off) % 16 == 0 and _addr(dH, dh_off) % 16 == 0
        r = np.zeros((), dtype=abi.HEAD_BWD_JOB_DT)
        r["dl"], r["W"], r["H"], r["dH"] = _addr(dl, dl_off), _addr(W, w_off), _addr(H, h_off), _addr(dH, dh_off)
        r["gW"], r["gb"] = _addr(gW, gw_off), _addr(gb, gb_off)
        r["rows"], r["n"], r["na"], r["ld_dl"] = rows, n, na, ld_dl
        self.rows.append(r)
        self.keep += [dl, W, H, dH, gW, gb]
        self.max_rows = max(self.max_rows, int(rows))
        self.dev_table = None
        return self

    def finalize(self):
        host = np.stack(self.rows)
        self.dev_table = torch.from_numpy(host.view(np.uint8).reshape(len(self.rows), -1).copy()).to(self.device)
        return self

    def launch(self, stream=None):
        if self.dev_table is None:
            self.finalize()
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        rc = abi.lib().fjsp_a2c_head_backward(C.c_void_p(self.dev_table.data_ptr()), len(self.rows), self.max_rows, C.c_void_p(st))
        if rc:
            abi.check(rc)
