#!/usr/bin/env python
"""Code generator for the per-instance EVOPF math (writes rpo_b200/csrc/evopf_gen.cuh).

The IEEE-14 network is fixed, so the sparsity pattern of the power-flow Jacobian, a fill-reducing
elimination order and every index map are known at build time.  This script performs the symbolic
work once (pattern, minimum-degree bus ordering, symbolic LU with fill) and emits straight-line
`__host__ __device__` code over locally-scoped arrays with literal indices, which nvcc scalarises
into registers: no pivot search, no index arithmetic, no zero multiplications, all network
constants folded into FFMA immediates.

Math (derivation in DESIGN.md §3; reference lines in rpo/env/electrical_grid/evopf.py):
  flows   P_i = vr_i a_i + vi_i b_i, Q_i = vi_i a_i - vr_i b_i, a + jb = Y (vr + j vi)   (:520-546)
  T1_ij = vr_i al_ij + vi_i be_ij, T2_ij = vi_i al_ij - vr_i be_ij, al + j be = Y_ij (vr_j + j vi_j)
  scaled flow Jacobian F (V-columns multiplied by V_j), i != j:
      dP_i/dth_j = T2_ij   V_j dP_i/dV_j = T1_ij   dQ_i/dth_j = -T1_ij   V_j dQ_i/dV_j = T2_ij
  i == j:  T2_ii - Q_i,    T1_ii + P_i,            P_i - T1_ii,          T2_ii + Q_i                (:614-661)
  reference Jacobian: J[:, va_j] = -F[:, th_j], J[:, vm_j] = -F[:, V_j] / V_j.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rpo_b200 import evopf_constants as K  # noqa: E402


def lit(x: float) -> str:
    """f32-rounded literal wrapped in RC() so the same header can be compiled in double for validation."""
    v = float(np.float32(x))
    return "RC(%.17g)" % v


class Gen(object):
    def __init__(self):
        self.G = K.G.astype(np.float32).astype(np.float64)
        self.B = K.B.astype(np.float32).astype(np.float64)
        nb = K.NBUS
        self.nb = nb
        self.nz = [[j for j in range(nb) if (self.G[i, j] != 0 or self.B[i, j] != 0)] for i in range(nb)]
        self.kidx = {}
        for i in range(nb):
            for j in self.nz[i]:
                self.kidx[(i, j)] = len(self.kidx)
        self.nnz = len(self.kidx)
        assert np.allclose(K.G, K.G.T) and np.allclose(K.B, K.B.T), "symmetric Y-bus expected"
        self.slack = int(K.SLACK[0])
        self.spv = [int(v) for v in K.SPV]
        self.pv = [int(v) for v in K.PV]
        self.pq = [int(v) for v in K.PQ]
        self.order = self.min_degree_order()
        self.unk, self.eqs = [], []
        for b in self.order:
            self.unk.append(("th", b))
            self.eqs.append(("P", b))
            if b in self.pq:
                self.unk.append(("V", b))
                self.eqs.append(("Q", b))
        self.n = len(self.unk)
        assert self.n == 22
        self.symbolic()
        self.out = []
        self.align_on = False
        self.align_cnt = 0

    # ---------------------------------------------------------------- symbolic work
    def min_degree_order(self):
        nodes = [b for b in range(self.nb) if b != self.slack]
        adj = {b: set(j for j in self.nz[b] if j != b and j != self.slack) for b in nodes}
        weight = {b: (2 if b in self.pq else 1) for b in nodes}
        order = []
        while adj:
            # minimise the fill-in flops of eliminating b: sum of neighbour weights (ties: lowest id)
            b = min(adj, key=lambda v: (sum(weight[u] for u in adj[v]), v))
            nb_ = adj.pop(b)
            for u in nb_:
                adj[u].discard(b)
                adj[u] |= (nb_ - {u})
            order.append(b)
        return order

    def symbolic(self):
        n = self.n
        pat = [[(self.unk[c][1] in self.nz[self.eqs[r][1]]) for c in range(n)] for r in range(n)]
        self.orig = [row[:] for row in pat]
        for k in range(n):
            assert pat[k][k]
            for i in range(k + 1, n):
                if pat[i][k]:
                    for j in range(k + 1, n):
                        if pat[k][j]:
                            pat[i][j] = True
        self.pat = pat
        self.luidx = {}
        for r in range(n):
            for c in range(n):
                if pat[r][c]:
                    self.luidx[(r, c)] = len(self.luidx)
        self.nlu = len(self.luidx)
        fl = 0
        for k in range(n):
            rows = [i for i in range(k + 1, n) if pat[i][k]]
            cols = [j for j in range(k + 1, n) if pat[k][j]]
            fl += len(rows) * (1 + len(cols))
        self.lu_ops = fl
        self.nL = sum(1 for (r, c) in self.luidx if r > c)
        self.nU = sum(1 for (r, c) in self.luidx if r <= c)

    # ---------------------------------------------------------------- expression helpers
    def F(self, eq, unk, pt="pt."):
        """C expression of the scaled flow-Jacobian entry F[eq, unk] or None if structurally zero."""
        (ek, i), (uk, j) = eq, unk
        if j not in self.nz[i]:
            return None
        k = self.kidx[(i, j)]
        if i != j:
            return {("P", "th"): "%sT2[%d]" % (pt, k), ("P", "V"): "%sT1[%d]" % (pt, k),
                    ("Q", "th"): "(-%sT1[%d])" % (pt, k), ("Q", "V"): "%sT2[%d]" % (pt, k)}[(ek, uk)]
        return {("P", "th"): "%sHd[%d]" % (pt, i), ("P", "V"): "%sNd[%d]" % (pt, i),
                ("Q", "th"): "%sMd[%d]" % (pt, i), ("Q", "V"): "%sLd[%d]" % (pt, i)}[(ek, uk)]

    def emit(self, s=""):
        self.out.append(s)
        if self.align_on and s.startswith("    ") and not s.startswith("     ") and s.rstrip().endswith(";"):
            self.align_cnt += 1
            if self.align_cnt % 32 == 0:
                self.out.append("    RPO_ALIGN(%d);" % (self.align_cnt // 32))

    def align(self, on):
        """Enable/disable instruction-fetch alignment markers (see RPO_ALIGN in rpo_math.cuh)."""
        self.align_on = on

    def accum(self, target, terms, init=None, negate=False):
        """Emit target = init (+|-) sum(a*b) as an FMA chain.  terms: list of (exprA, exprB)."""
        sign = "-" if negate else ""
        if init is None:
            if not terms:
                self.emit("    %s = RC(0.0);" % target)
                return
            a, b = terms[0]
            self.emit("    %s = %s(%s) * (%s);" % (target, sign, a, b))
            terms = terms[1:]
        else:
            self.emit("    %s = %s;" % (target, init))
        for a, b in terms:
            self.emit("    %s = FMA(%s(%s), (%s), %s);" % (target, sign, a, b, target))

    # ---------------------------------------------------------------- emitters
    def gen_header(self):
        e = self.emit
        e("// GENERATED by tools/gen_evopf.py — do not edit by hand.")
        e("// Per-instance EVOPF-v0 math over locally scoped arrays with literal indices (see DESIGN.md §3).")
        e("// elimination order (bus ids, 0-based): %s" % self.order)
        e("// nnz(Y) = %d, Newton n = %d, nnz(L) = %d, nnz(U) = %d (original nnz(J_N) = %d), LU mul/fma ops = %d"
          % (self.nnz, self.n, self.nL, self.nU, sum(sum(r) for r in self.orig), self.lu_ops))
        e("#pragma once")
        e("#define EVOPF_NNZY %d" % self.nnz)
        e("#define EVOPF_NN %d" % self.n)
        e("#define EVOPF_NLU %d" % self.nlu)
        e("")
        e("struct EvopfPt {")
        e("    real vr[14], vi[14];      // rectangular voltages")
        e("    real P[14], Q[14];        // bus power flows")
        e("    real T1[EVOPF_NNZY], T2[EVOPF_NNZY];")
        e("    real Hd[14], Nd[14], Md[14], Ld[14];  // diagonal entries of the scaled flow Jacobian")
        e("};")
        e("")

    def gen_eval(self):
        """evopf_eval_{flows,full}: rectangular voltages, bus flows and (full) the Jacobian terms.

        Flows use the reference's association.  The Jacobian terms use the branch-pair form: per undirected branch
        {i<j} with c = vr_i vr_j + vi_i vi_j (= V_i V_j cos th_ij) and s = vi_i vr_j - vr_i vi_j (= V_i V_j sin th_ij):
        T1_ij = G c + B s, T1_ji = G c - B s, T2_ij = G s - B c, T2_ji = -G s - B c (Y symmetric), and on the
        diagonal T1_ii = G_ii V_i^2, T2_ii = -B_ii V_i^2."""
        e = self.emit
        edges = [(i, j) for i in range(self.nb) for j in self.nz[i] if i < j]
        for full in (0, 1):
            name = "evopf_eval_full" if full else "evopf_eval_flows"
            e("// y -> point quantities (%s)" % ("flows + Jacobian terms" if full else "flows only"))
            e("RPO_HD void %s(const real* y, EvopfPt& pt) {" % name)
            self.align(bool(full))
            # voltage angles are small (|va| < 1 rad) in every realistic state: skip the range reduction then
            e("    real amax_ = FABS(y[%d]);" % K.VA0)
            for i in range(1, self.nb):
                e("    amax_ = FMAX(amax_, FABS(y[%d]));" % (K.VA0 + i))
            e("    if (amax_ <= RC(1.6)) {")
            for i in range(self.nb):
                e("        { real s_, c_; SINCOS_SMALL(y[%d], &s_, &c_); pt.vr[%d] = y[%d] * c_; pt.vi[%d] = y[%d] * s_; }"
                  % (K.VA0 + i, i, K.VM0 + i, i, K.VM0 + i))
            e("    } else {")
            for i in range(self.nb):
                e("        { real s_, c_; SINCOS(y[%d], &s_, &c_); pt.vr[%d] = y[%d] * c_; pt.vi[%d] = y[%d] * s_; }"
                  % (K.VA0 + i, i, K.VM0 + i, i, K.VM0 + i))
            e("    }")
            # flows in the reference's association (a + jb = Y v, then P = vr a + vi b, Q = vi a - vr b; evopf.py:527-538):
            # measured 20 % less residual noise than summing the per-branch T terms, and the residual is what the
            # Newton tolerance is judged on
            for i in range(self.nb):
                terms_a, terms_b = [], []
                for j in self.nz[i]:
                    g, b = self.G[i, j], self.B[i, j]
                    if g != 0:
                        terms_a.append((lit(g), "pt.vr[%d]" % j))
                        terms_b.append((lit(g), "pt.vi[%d]" % j))
                    terms_a.append((lit(-b), "pt.vi[%d]" % j))
                    terms_b.append((lit(b), "pt.vr[%d]" % j))
                e("    {")
                e("    real a_, b_;")
                self.accum("a_", terms_a)
                self.accum("b_", terms_b)
                e("    pt.P[%d] = FMA(pt.vr[%d], a_, pt.vi[%d] * b_);" % (i, i, i))
                e("    pt.Q[%d] = FMA(pt.vi[%d], a_, -(pt.vr[%d] * b_));" % (i, i, i))
                e("    }")
            if full:
                # Jacobian terms per undirected branch (accuracy-uncritical): 10 ops per branch for 4 numbers x 2 directions
                for (i, j) in edges:
                    g, b = self.G[i, j], self.B[i, j]
                    kij, kji = self.kidx[(i, j)], self.kidx[(j, i)]
                    e("    {")
                    e("    const real c_ = FMA(pt.vr[%d], pt.vr[%d], pt.vi[%d] * pt.vi[%d]);" % (i, j, i, j))
                    e("    const real s_ = FMA(pt.vi[%d], pt.vr[%d], -(pt.vr[%d] * pt.vi[%d]));" % (i, j, i, j))
                    e("    const real bs_ = %s * s_; const real bc_ = %s * c_;" % (lit(b), lit(b)))
                    if g != 0:
                        e("    pt.T1[%d] = FMA(%s, c_, bs_); pt.T1[%d] = FMA(%s, c_, -bs_);" % (kij, lit(g), kji, lit(g)))
                        e("    pt.T2[%d] = FMA(%s, s_, -bc_); pt.T2[%d] = FMA(%s, s_, -bc_);" % (kij, lit(g), kji, lit(-g)))
                    else:
                        e("    pt.T1[%d] = bs_; pt.T1[%d] = -bs_;" % (kij, kji))
                        e("    pt.T2[%d] = -bc_; pt.T2[%d] = -bc_;" % (kij, kji))
                    e("    }")
                for i in range(self.nb):
                    g, b = self.G[i, i], self.B[i, i]
                    e("    {")
                    e("    const real v2_ = y[%d] * y[%d];" % (K.VM0 + i, K.VM0 + i))
                    e("    const real t1d_ = %s * v2_, t2d_ = %s * v2_;" % (lit(g), lit(-b)))
                    e("    pt.Hd[%d] = t2d_ - pt.Q[%d]; pt.Nd[%d] = t1d_ + pt.P[%d];" % (i, i, i, i))
                    e("    pt.Md[%d] = pt.P[%d] - t1d_; pt.Ld[%d] = t2d_ + pt.Q[%d];" % (i, i, i, i))
                    e("    }")
            self.align(False)
            e("}")
            e("")

    def gen_factor(self):
        e = self.emit
        n = self.n
        e("// In-place sparse LU (no pivoting, static order) of the scaled Newton matrix A~ = F[N rows, w cols].")
        e("// lu[] holds the strict-lower multipliers and U; ipiv[] the pivot reciprocals.  Returns min |pivot|.")
        e("RPO_HD real evopf_factor(const EvopfPt& pt, real* lu, real* ipiv) {")
        self.align(True)
        zero = set()
        for (r, c), idx in sorted(self.luidx.items(), key=lambda kv: kv[1]):
            if self.orig[r][c]:
                e("    lu[%d] = %s;" % (idx, self.F(self.eqs[r], self.unk[c])))
            else:
                zero.add((r, c))
        e("    real minpiv = RC(3.0e38);")
        for k in range(n):
            kk = self.luidx[(k, k)]
            e("    minpiv = FMIN(minpiv, FABS(lu[%d]));" % kk)
            e("    ipiv[%d] = RCP(lu[%d]);" % (k, kk))
            rows = [i for i in range(k + 1, n) if self.pat[i][k]]
            cols = [j for j in range(k + 1, n) if self.pat[k][j]]
            for i in rows:
                ik = self.luidx[(i, k)]
                assert (i, k) not in zero
                e("    lu[%d] *= ipiv[%d];" % (ik, k))
                for j in cols:
                    ij, kj = self.luidx[(i, j)], self.luidx[(k, j)]
                    assert (k, j) not in zero
                    if (i, j) in zero:
                        e("    lu[%d] = -(lu[%d] * lu[%d]);" % (ij, ik, kj))
                        zero.discard((i, j))
                    else:
                        e("    lu[%d] = FMA(-lu[%d], lu[%d], lu[%d]);" % (ij, ik, kj, ij))
        assert not zero
        self.align(False)
        e("    return minpiv;")
        e("}")
        e("")
        # solves
        e("// x <- A~^{-1} x  (x indexed by elimination position: in = equations, out = unknowns)")
        e("RPO_HD void evopf_solve(const real* lu, const real* ipiv, real* x) {")
        self.align(True)
        for k in range(n):
            for i in range(k + 1, n):
                if self.pat[i][k]:
                    e("    x[%d] = FMA(-lu[%d], x[%d], x[%d]);" % (i, self.luidx[(i, k)], k, i))
        for k in reversed(range(n)):
            for j in range(k + 1, n):
                if self.pat[k][j]:
                    e("    x[%d] = FMA(-lu[%d], x[%d], x[%d]);" % (k, self.luidx[(k, j)], j, k))
            e("    x[%d] *= ipiv[%d];" % (k, k))
        self.align(False)
        e("}")
        e("")
        e("// x <- A~^{-T} x  (in = unknown positions, out = equation positions)")
        e("RPO_HD void evopf_solve_t(const real* lu, const real* ipiv, real* x) {")
        self.align(True)
        # U^T w = b : forward over columns of U
        for k in range(n):
            for j in range(k):
                if self.pat[j][k]:
                    e("    x[%d] = FMA(-lu[%d], x[%d], x[%d]);" % (k, self.luidx[(j, k)], j, k))
            e("    x[%d] *= ipiv[%d];" % (k, k))
        # L^T x = w : backward, unit diagonal
        for k in reversed(range(n)):
            for i in range(k + 1, n):
                if self.pat[i][k]:
                    e("    x[%d] = FMA(-lu[%d], x[%d], x[%d]);" % (k, self.luidx[(i, k)], i, k))
        self.align(False)
        e("}")
        e("")

    def eq_pos(self, kind, bus):
        return self.eqs.index((kind, bus))

    def unk_pos(self, kind, bus):
        return self.unk.index((kind, bus))

    def gen_newton(self):
        e = self.emit
        e("// Newton right-hand side r[k] = injection - flow of equation k (elimination order), from an evaluated point.")
        e("RPO_HD void evopf_newton_rhs(const real* x, const real* y, const EvopfPt& pt, real* r) {")
        for k, (kind, i) in enumerate(self.eqs):
            if kind == "P":
                if i in self.pv:
                    g = self.spv.index(i)
                    e("    r[%d] = ((y[%d] + y[%d]) - x[%d]) - pt.P[%d];" % (k, K.PG0 + g, K.PE0 + g, i, i))
                else:
                    e("    r[%d] = (-x[%d]) - pt.P[%d];" % (k, i, i))
            else:
                e("    r[%d] = (-x[%d]) - pt.Q[%d];" % (k, 14 + i, i))
        e("}")
        e("")
        e("// Apply the solved (scaled) Newton step to y; returns ||delta||_2^2 of the unscaled update.")
        e("RPO_HD real evopf_newton_update(const real* r, real* y) {")
        e("    real n2 = RC(0.0);")
        for k, (kind, j) in enumerate(self.unk):
            if kind == "th":
                e("    y[%d] += r[%d]; n2 = FMA(r[%d], r[%d], n2);" % (K.VA0 + j, k, k, k))
            else:
                e("    { const real dv_ = y[%d] * r[%d]; y[%d] += dv_; n2 = FMA(dv_, dv_, n2); }" % (K.VM0 + j, k, K.VM0 + j))
        e("    return n2;")
        e("}")
        e("")
        e("// One Newton iteration of the completion (PFFunctionFn.forward evopf.py:826-835): y's Newton unknowns are")
        e("// updated in place; returns ||delta||_2^2 of the *unscaled* update.  pd/qd = x[0:14], x[14:28].")
        e("RPO_HD real evopf_newton_iter(const real* x, real* y, EvopfPt& pt, real* minpiv) {")
        e("    real lu[EVOPF_NLU], ipiv[EVOPF_NN], r[EVOPF_NN];")
        e("    evopf_eval_full(y, pt);")
        e("    evopf_newton_rhs(x, y, pt, r);")
        e("    *minpiv = evopf_factor(pt, lu, ipiv);")
        e("    evopf_solve(lu, ipiv, r);")
        e("    return evopf_newton_update(r, y);")
        e("}")
        e("")
        e("// Step 0 of the completion: scatter z, flat start (evopf.py:794-807).")
        e("RPO_HD void evopf_init(const real* z, real* y) {")
        for k in range(K.YDIM):
            e_ = None
        for g in range(K.NG):
            e("    y[%d] = RC(0.0); y[%d] = RC(0.0);" % (K.PG0 + g, K.QG0 + g))
        for k, g in enumerate(K.PV_):
            e("    y[%d] = z[%d];" % (K.PG0 + int(g), int(K.PG_PV_ZIDX[k])))
        for k, b in enumerate(K.SPV):
            e("    y[%d] = z[%d];" % (K.VM0 + int(b), int(K.VM_SPV_ZIDX[k])))
        for b in self.pq:
            e("    y[%d] = %s;" % (K.VM0 + b, lit(K.VM_INIT[b])))
        e("    y[%d] = %s;" % (K.VA0 + self.slack, lit(K.SLACK_VA[0])))
        for b in self.pv + self.pq:
            e("    y[%d] = %s;" % (K.VA0 + b, lit(K.VA_INIT[b])))
        for g in range(K.NE):
            e("    y[%d] = z[%d];" % (K.PE0 + g, int(K.PE_SPV_ZIDX[g])))
        e("}")
        e("")
        e("// Step 2 of the completion: qg at all gens, pg at the slack (evopf.py:844-848), from flows at the final point.")
        e("RPO_HD void evopf_tail(const real* x, real* y, const EvopfPt& pt) {")
        for g, b in enumerate(self.spv):
            e("    y[%d] = x[%d] + pt.Q[%d];" % (K.QG0 + g, 14 + b, b))
        e("    y[%d] = (x[%d] - y[%d]) + pt.P[%d];" % (K.PG0 + 0, self.slack, K.PE0 + 0, self.slack))
        e("}")
        e("")

    def gen_refine(self):
        """float64 flows for the mixed-precision refinement of the completion: the float32 sparse LU of the
        last Newton iteration is reused for a chord step whose right-hand side is the residual with the flows accumulated
        in float64 (the FP64 pipe is otherwise idle), which removes the float32 evaluation noise from the converged point;
        the closed-form tail (qg, slack pg) is taken from the same float64 flows plus the first-order effect of the step."""
        e = self.emit

        def dlit(x):
            return "%.17g" % float(np.float32(x))

        e("// Bus power flows accumulated in float64 (network constants = the float32-rounded Y-bus the reference holds,")
        e("// evopf.py:260-266); sincos by rpo_sincos_mixed (SINCOS_REFINE).")
        e("RPO_HD void evopf_flows_f64(const real* y, double* P, double* Q) {")
        e("    double vr[14], vi[14];")
        for i in range(self.nb):
            e("    { double s_, c_; SINCOS_REFINE(y[%d], &s_, &c_); vr[%d] = (double)y[%d] * c_; vi[%d] = (double)y[%d] * s_; }"
              % (K.VA0 + i, i, K.VM0 + i, i, K.VM0 + i))
        for i in range(self.nb):
            ta, tb = [], []
            for j in self.nz[i]:
                g, b = self.G[i, j], self.B[i, j]
                if g != 0:
                    ta.append((dlit(g), "vr[%d]" % j))
                    tb.append((dlit(g), "vi[%d]" % j))
                ta.append((dlit(-b), "vi[%d]" % j))
                tb.append((dlit(b), "vr[%d]" % j))
            e("    {")
            e("    double a_ = %s * %s, b_ = %s * %s;" % (ta[0][0], ta[0][1], tb[0][0], tb[0][1]))
            for c, v in ta[1:]:
                e("    a_ = fma(%s, %s, a_);" % (c, v))
            for c, v in tb[1:]:
                e("    b_ = fma(%s, %s, b_);" % (c, v))
            # every bus needs P except none; Q is needed at the PQ buses (Newton rows) and the generator buses (tail)
            e("    P[%d] = fma(vr[%d], a_, vi[%d] * b_);" % (i, i, i))
            e("    Q[%d] = fma(vi[%d], a_, -(vr[%d] * b_));" % (i, i, i))
            e("    }")
        e("}")
        e("")
        e("// Newton right-hand side from float64 flows (rounded to `real` once, at the end).")
        e("RPO_HD void evopf_newton_rhs_f64(const real* x, const real* y, const double* P, const double* Q, real* r) {")
        for k, (kind, i) in enumerate(self.eqs):
            if kind == "P":
                if i in self.pv:
                    g = self.spv.index(i)
                    e("    r[%d] = (real)((((double)y[%d] + (double)y[%d]) - (double)x[%d]) - P[%d]);" % (k, K.PG0 + g, K.PE0 + g, i, i))
                else:
                    e("    r[%d] = (real)((-(double)x[%d]) - P[%d]);" % (k, i, i))
            else:
                e("    r[%d] = (real)((-(double)x[%d]) - Q[%d]);" % (k, 14 + i, i))
        e("}")
        e("")
        e("// Step 2 of the completion (evopf.py:844-848) after a refinement step: P, Q are the float64 flows at the point the")
        e("// step started from, r the scaled step just applied (theta += r, V += V r), pt the Jacobian terms of the last Newton")
        e("// point: flow(y_new) = flow(y_old) + F[L, w] r up to |dF| |r| <= ~40 * RPO_REFINE_SWITCH * tol (4e-7 at 1e-3 / 1e-5).")
        e("RPO_HD void evopf_tail_refined(const real* x, real* y, const EvopfPt& pt, const double* P, const double* Q, const real* r) {")
        Lrows = self.L_rows()
        for li, row in enumerate(Lrows):
            terms = []
            for c, unk in enumerate(self.unk):
                f = self.F(row, unk)
                if f is not None:
                    terms.append((f, "r[%d]" % c))
            e("    {")
            e("    real u_;")
            self.accum("u_", terms)
            if row[0] == "P":
                e("    y[%d] = (real)((((double)x[%d] - (double)y[%d]) + P[%d]) + (double)u_);" % (K.PG0 + 0, self.slack, K.PE0 + 0, self.slack))
            else:
                g = self.spv.index(row[1])
                e("    y[%d] = (real)(((double)x[%d] + Q[%d]) + (double)u_);" % (K.QG0 + g, 14 + row[1], row[1]))
            e("    }")
        e("}")
        e("")

    def gen_resid(self):
        e = self.emit
        e("// eq_resid (evopf.py:520-546) from a flows-evaluated point; returns max |resid|, optionally stores the 28 values.")
        e("RPO_HD real evopf_eq_resid(const real* x, const real* y, const EvopfPt& pt, real* out) {")
        e("    real m = RC(0.0), r_;")
        for i in range(self.nb):
            if i in self.spv:
                g = self.spv.index(i)
                e("    r_ = ((y[%d] + y[%d]) - x[%d]) - pt.P[%d];" % (K.PG0 + g, K.PE0 + g, i, i))
            else:
                e("    r_ = (-x[%d]) - pt.P[%d];" % (i, i))
            e("    if (out) out[%d] = r_; m = FMAX(m, FABS(r_));" % i)
        for i in range(self.nb):
            if i in self.spv:
                g = self.spv.index(i)
                e("    r_ = (y[%d] - x[%d]) - pt.Q[%d];" % (K.QG0 + g, 14 + i, i))
            else:
                e("    r_ = (-x[%d]) - pt.Q[%d];" % (14 + i, i))
            e("    if (out) out[%d] = r_; m = FMAX(m, FABS(r_));" % (14 + i))
        e("    return m;")
        e("}")
        e("")
        e("// Battery bounds in float64 then rounded, as Battery.ineq_resid does in numpy (evopf.py:109-116).")
        e("// One battery (the same numbers are the volatile action-box rows: Battery.update_bound evopf.py:127-143).")
        e("RPO_HD void evopf_pe_bound1(real soc, real* pe_max, real* pe_min) {")
        e("    const double s = (double)soc;")
        e("    const double hi = (1.0 / %r) * fmin(%r, %r - s);" % (K.ETA_IN, K.B_PMAX, K.B_HIGH))
        e("    const double lo = %r * fmax(%r, %r - s);" % (K.ETA_OUT, K.B_PMIN, K.B_LOW))
        e("    *pe_max = (real)hi; *pe_min = (real)lo;")
        e("}")
        e("RPO_HD void evopf_pe_bounds(const real* soc, real* pe_max, real* pe_min) {")
        e("    for (int g = 0; g < 5; ++g) evopf_pe_bound1(soc[g], pe_max + g, pe_min + g);")
        e("}")
        e("")
        e("// ineq_resid (evopf.py:548-563): 58 rows; returns max ineq_dist; optionally stores the residuals;")
        e("// d[43] receives ineq_jac^T . 1[dist > 0] (evopf.py:590-594) when non-null.  pe_max/pe_min from evopf_pe_bounds.")
        e("RPO_HD real evopf_ineq(const real* pe_max, const real* pe_min, const real* y, real* out, real* d) {")
        e("    real m = RC(0.0), hi_, lo_;")
        rows = []

        def pair(yidx, hi, lo, row_hi, row_lo):
            e("    hi_ = y[%d] - %s; lo_ = %s - y[%d];" % (yidx, hi, lo, yidx))
            e("    if (out) { out[%d] = hi_; out[%d] = lo_; }" % (row_hi, row_lo))
            e("    m = FMAX(m, FMAX(hi_, lo_));")
            e("    if (d) d[%d] = (hi_ > RC(0.0) ? RC(1.0) : RC(0.0)) - (lo_ > RC(0.0) ? RC(1.0) : RC(0.0));" % yidx)
        ng, nb = K.NG, K.NBUS
        for g in range(ng):
            pair(K.PG0 + g, lit(K.PMAX_PU[g]), lit(K.PMIN_PU[g]), g, ng + g)
        for g in range(ng):
            pair(K.QG0 + g, lit(K.QMAX_PU[g]), lit(K.QMIN_PU[g]), 2 * ng + g, 3 * ng + g)
        for i in range(nb):
            pair(K.VM0 + i, lit(K.VMAX_PU[i]), lit(K.VMIN_PU[i]), 4 * ng + i, 4 * ng + nb + i)
        for g in range(ng):
            pair(K.PE0 + g, "pe_max[%d]" % g, "pe_min[%d]" % g, 4 * ng + 2 * nb + g, 5 * ng + 2 * nb + g)
        e("    if (d) {")
        for i in range(nb):
            e("        d[%d] = RC(0.0);" % (K.VA0 + i))
        e("    }")
        e("    return m;")
        e("}")
        e("")

    def L_rows(self):
        return [("P", self.slack)] + [("Q", b) for b in self.spv]

    def gen_transposed_rhs(self, fn_name, doc):
        """rhs[c] = -(D gvec_w)[c] - sum_{rows in L} F[row, unk c] * tL[row];  tL: P_slack, Q_spv[0..4]."""
        e = self.emit
        e("// %s" % doc)
        e("RPO_HD void %s(const real* y, const EvopfPt& pt, const real* gvec, const real* tL, real* rhs) {" % fn_name)
        Lrows = self.L_rows()
        for c, (kind, j) in enumerate(self.unk):
            terms = []
            for li, row in enumerate(Lrows):
                f = self.F(row, (kind, j))
                if f is not None:
                    terms.append((f, "tL[%d]" % li))
            if kind == "th":
                init = "-gvec[%d]" % (K.VA0 + j)
            else:
                init = "-(y[%d] * gvec[%d])" % (K.VM0 + j, K.VM0 + j)
            self.accum("rhs[%d]" % c, terms, init=init, negate=True)
        e("}")
        e("")

    def gen_colsums(self):
        """For columns V_j (j in spv) and th_slack: sum over all rows of F[row, col] * lam[row], lam = (lamP[14], lamQ[14])."""
        e = self.emit
        e("// cs[g] = sum_rows F[row, V_spv[g]] lam[row] (g = 0..4);  cs[5] = sum_rows F[row, th_slack] lam[row].")
        e("RPO_HD void evopf_colsums(const EvopfPt& pt, const real* lamP, const real* lamQ, real* cs) {")
        cols = [("V", b) for b in self.spv] + [("th", self.slack)]
        for ci, col in enumerate(cols):
            terms = []
            for i in range(self.nb):
                fp = self.F(("P", i), col)
                if fp is not None:
                    terms.append((fp, "lamP[%d]" % i))
                fq = self.F(("Q", i), col)
                if fq is not None:
                    terms.append((fq, "lamQ[%d]" % i))
            self.accum("cs[%d]" % ci, terms)
        e("}")
        e("")

    def gen_backward(self):
        e = self.emit
        self.gen_transposed_rhs(
            "evopf_rhs_t",
            "Right-hand side of the transposed Newton solve shared by the GRG backward and the correction direction:"
            " A~^T lam_N = -D g_w - F[L,w]^T t_L.")
        self.gen_colsums()
        e("// GRG backward (PFFunctionFn.backward evopf.py:857-910) with J recomputed at y.  g = dL/dy (43).")
        e("// dz (14) and dx (28, may be null).")
        e("RPO_HD real evopf_backward(const real* y, const real* g, real* dz, real* dx) {")
        e("    EvopfPt pt; real lu[EVOPF_NLU], ipiv[EVOPF_NN], rhs[EVOPF_NN], tL[6], lamP[14], lamQ[14], cs[6];")
        e("    evopf_eval_full(y, pt);")
        e("    const real minpiv = evopf_factor(pt, lu, ipiv);")
        e("    tL[0] = g[%d];" % (K.PG0 + 0))
        for gi in range(K.NG):
            e("    tL[%d] = g[%d];" % (1 + gi, K.QG0 + gi))
        e("    evopf_rhs_t(y, pt, g, tL, rhs);")
        e("    evopf_solve_t(lu, ipiv, rhs);")
        e("    lamP[%d] = tL[0];" % self.slack)
        for gi, b in enumerate(self.spv):
            e("    lamQ[%d] = tL[%d];" % (b, 1 + gi))
        for k, (kind, i) in enumerate(self.eqs):
            e("    lam%s[%d] = rhs[%d];" % (kind, i, k))
        e("    evopf_colsums(pt, lamP, lamQ, cs);")
        for k, gidx in enumerate(K.PV_):
            b = self.spv[int(gidx)]
            e("    dz[%d] = g[%d] - lamP[%d];" % (int(K.PG_PV_ZIDX[k]), K.PG0 + int(gidx), b))
        for gi, b in enumerate(self.spv):
            e("    dz[%d] = FMA(cs[%d], RCP(y[%d]), g[%d]);" % (int(K.VM_SPV_ZIDX[gi]), gi, K.VM0 + b, K.VM0 + b))
        for gi, b in enumerate(self.spv):
            e("    dz[%d] = g[%d] + lamP[%d];" % (int(K.PE_SPV_ZIDX[gi]), K.PE0 + gi, b))
        e("    if (dx) {")
        for i in range(self.nb):
            e("        dx[%d] = lamP[%d]; dx[%d] = lamQ[%d];" % (i, i, 14 + i, i))
        e("    }")
        e("    return minpiv;")
        e("}")
        e("")


    def gen_vjp(self):
        """Vector-Jacobian products of the two residual functions w.r.t. the action (what torch autograd computes for
        the reference's plain-torch eq_resid / ineq_resid, evopf.py:520-563) and the GRG backward with the inequality
        penalty folded in."""
        e = self.emit
        e("// dy += (d eq_resid / d y)^T [gP | gQ]  — the TRUE derivative of evopf.py:520-546 (autograd of the torch code):")
        e("// +I for pe on the P rows (eq_resid adds pe, :532; eq_jac's -I at :639-640 is NOT what autograd sees).")
        e("// pt must hold evopf_eval_full(y).")
        e("RPO_HD void evopf_eq_vjp(const real* y, const EvopfPt& pt, const real* gP, const real* gQ, real* dy) {")
        for g, b in enumerate(self.spv):
            e("    dy[%d] += gP[%d]; dy[%d] += gQ[%d]; dy[%d] += gP[%d];" % (K.PG0 + g, b, K.QG0 + g, b, K.PE0 + g, b))
        for j in range(self.nb):
            for kind in ("th", "V"):
                terms = []
                for i in range(self.nb):
                    fp = self.F(("P", i), (kind, j))
                    if fp is not None:
                        terms.append((fp, "gP[%d]" % i))
                    fq = self.F(("Q", i), (kind, j))
                    if fq is not None:
                        terms.append((fq, "gQ[%d]" % i))
                e("    {")
                e("    real cs_;")
                self.accum("cs_", terms)
                if kind == "th":
                    e("    dy[%d] -= cs_;" % (K.VA0 + j))
                else:
                    e("    dy[%d] -= cs_ / y[%d];" % (K.VM0 + j, K.VM0 + j))
                e("    }")
        e("}")
        e("")
        e("// dy += (d ineq_resid / d y)^T (g . m): the constant +-1 selector of evopf.py:663-707 applied to the 58 row")
        e("// gradients g; relu != 0 masks rows by [resid >= 0] (the gradient torch.clamp(resid, 0) passes, evopf.py:565-567).")
        e("// `scale` multiplies g (per-row weight of a shared dual vector).")
        e("RPO_HD void evopf_ineq_vjp(const real* pe_max, const real* pe_min, const real* y, const real* g, real scale, int relu, real* dy) {")
        e("    real hi_, lo_;")

        def pair(yidx, hi, lo, row_hi, row_lo):
            e("    hi_ = y[%d] - %s; lo_ = %s - y[%d];" % (yidx, hi, lo, yidx))
            e("    dy[%d] += scale * (((!relu || hi_ >= RC(0.0)) ? g[%d] : RC(0.0)) - ((!relu || lo_ >= RC(0.0)) ? g[%d] : RC(0.0)));"
              % (yidx, row_hi, row_lo))
        ng, nb = K.NG, K.NBUS
        for g in range(ng):
            pair(K.PG0 + g, lit(K.PMAX_PU[g]), lit(K.PMIN_PU[g]), g, ng + g)
        for g in range(ng):
            pair(K.QG0 + g, lit(K.QMAX_PU[g]), lit(K.QMIN_PU[g]), 2 * ng + g, 3 * ng + g)
        for i in range(nb):
            pair(K.VM0 + i, lit(K.VMAX_PU[i]), lit(K.VMIN_PU[i]), 4 * ng + i, 4 * ng + nb + i)
        for g in range(ng):
            pair(K.PE0 + g, "pe_max[%d]" % g, "pe_min[%d]" % g, 4 * ng + 2 * nb + g, 5 * ng + 2 * nb + g)
        e("}")
        e("")
    def gen_direction(self):
        e = self.emit
        e("// Reduced correction direction = EVOPFEnv.ineq_partial_grad (evopf.py:596-612) in two-solve form")
        e("// (SURVEY App. B.4) on the 22x22 Newton block (28x28 J_O is block triangular, SURVEY H3).")
        e("// Also evaluates the loop test of grad_steps (rpo_ddpg.py:271-272): *eqmax = max|eq_resid|, *ineqmax = max ineq_dist.")
        e("// post-factor part of the direction: d = sign-of-violation vector (43), lu/ipiv = factors of the Newton block at y.")
        e("// Split around its plain solve so that the fused projection loop can share ONE inlined copy of evopf_solve between")
        e("// the Newton iteration and the correction step:  pre (transposed solve, g_P, rhs = -q_N) -> evopf_solve -> fin.")
        e("// keep[22] = [q_L(6) | gpg(5) | gpe(5) | gvm(5) | gva0] carries the values fin needs across the solve.")
        e("#define EVOPF_DIRKEEP 22")
        e("RPO_HD void evopf_direction_pre(const real* y, const EvopfPt& pt, const real* lu, const real* ipiv, const real* d, real* rhs, real* keep) {")
        e("    real tL[6], tP[14], tQ[14], cs[6], q[28];")
        e("    tL[0] = d[%d];" % (K.PG0 + 0))
        for gi in range(K.NG):
            e("    tL[%d] = d[%d];" % (1 + gi, K.QG0 + gi))
        e("    evopf_rhs_t(y, pt, d, tL, rhs);")
        e("    evopf_solve_t(lu, ipiv, rhs);")
        e("    tP[%d] = tL[0];" % self.slack)
        for gi, b in enumerate(self.spv):
            e("    tQ[%d] = tL[%d];" % (b, 1 + gi))
        for k, (kind, i) in enumerate(self.eqs):
            e("    t%s[%d] = rhs[%d];" % (kind, i, k))
        e("    evopf_colsums(pt, tP, tQ, cs);")
        # g_P
        e("    real gpg[5], gpe[5], h[5], gvm[5];")
        e("    gpg[0] = RC(0.0);")
        for gidx in K.PV_:
            gidx = int(gidx)
            e("    gpg[%d] = d[%d] - tP[%d];" % (gidx, K.PG0 + gidx, self.spv[gidx]))
        for gi, b in enumerate(self.spv):
            e("    gpe[%d] = d[%d] + tP[%d];" % (gi, K.PE0 + gi, b))
        for gi, b in enumerate(self.spv):
            e("    { const real iv_ = RCP(y[%d]); gvm[%d] = FMA(cs[%d], iv_, d[%d]); h[%d] = gvm[%d] * iv_; }"
              % (K.VM0 + b, gi, gi, K.VM0 + b, gi, gi))
        e("    const real gva0 = cs[5];")
        # q = J_P g_P (28 over equations), then rhs2 = -q_N
        for i in range(self.nb):
            for kind in ("P", "Q"):
                terms = []
                for gi, b in enumerate(self.spv):
                    f = self.F((kind, i), ("V", b))
                    if f is not None:
                        terms.append((f, "h[%d]" % gi))
                f = self.F((kind, i), ("th", self.slack))
                if f is not None:
                    terms.append((f, "gva0"))
                init = None
                if kind == "P" and i in self.spv:
                    gi = self.spv.index(i)
                    init = "(gpg[%d] - gpe[%d])" % (gi, gi)
                tgt = "q[%d]" % (i if kind == "P" else 14 + i)
                if init is None and not terms:
                    e("    %s = RC(0.0);" % tgt)
                elif init is None:
                    self.accum(tgt, terms, negate=True)
                else:
                    self.accum(tgt, terms, init=init, negate=True)
        for k, (kind, i) in enumerate(self.eqs):
            e("    rhs[%d] = -q[%d];" % (k, i if kind == "P" else 14 + i))
        Lrows = self.L_rows()
        for li, row in enumerate(Lrows):
            e("    keep[%d] = q[%d];" % (li, row[1] if row[0] == "P" else 14 + row[1]))
        for gi in range(5):
            e("    keep[%d] = gpg[%d]; keep[%d] = gpe[%d]; keep[%d] = gvm[%d];" % (6 + gi, gi, 11 + gi, gi, 16 + gi, gi))
        e("    keep[21] = gva0;")
        e("}")
        e("")
        e("RPO_HD void evopf_direction_fin(const real* y, const EvopfPt& pt, const real* rhs, const real* keep, real* out) {")
        # u_s = q_L + F[L, w] u~ ; out_O = -u
        Lout = [K.PG0 + 0] + [K.QG0 + gi for gi in range(K.NG)]
        assert len(Lrows) == 6
        for li, row in enumerate(Lrows):
            terms = []
            for c, unk in enumerate(self.unk):
                f = self.F(row, unk)
                if f is not None:
                    terms.append((f, "rhs[%d]" % c))
            e("    {")
            e("    real u_;")
            self.accum("u_", terms, init="keep[%d]" % li)
            e("    out[%d] = -u_;" % Lout[li])
            e("    }")
        for gidx in K.PV_:
            e("    out[%d] = keep[%d];" % (K.PG0 + int(gidx), 6 + int(gidx)))
        for gi, b in enumerate(self.spv):
            e("    out[%d] = keep[%d];" % (K.VM0 + b, 16 + gi))
        e("    out[%d] = keep[21];" % (K.VA0 + self.slack))
        for gi in range(K.NE):
            e("    out[%d] = keep[%d];" % (K.PE0 + gi, 11 + gi))
        for c, (kind, j) in enumerate(self.unk):
            if kind == "th":
                e("    out[%d] = -rhs[%d];" % (K.VA0 + j, c))
            else:
                e("    out[%d] = -(y[%d] * rhs[%d]);" % (K.VM0 + j, K.VM0 + j, c))
        e("}")
        e("")
        e("RPO_HD void evopf_direction_post(const real* y, const EvopfPt& pt, const real* lu, const real* ipiv, const real* d, real* out) {")
        e("    real rhs[EVOPF_NN], keep[EVOPF_DIRKEEP];")
        e("    evopf_direction_pre(y, pt, lu, ipiv, d, rhs, keep);")
        e("    evopf_solve(lu, ipiv, rhs);")
        e("    evopf_direction_fin(y, pt, rhs, keep, out);")
        e("}")
        e("")
        e("RPO_HD real evopf_direction(const real* x, const real* pe_max, const real* pe_min, const real* y, real* out, real* eqmax, real* ineqmax) {")
        e("    EvopfPt pt; real lu[EVOPF_NLU], ipiv[EVOPF_NN], d[43];")
        e("    evopf_eval_full(y, pt);")
        e("    *eqmax = evopf_eq_resid(x, y, pt, (real*)0);")
        e("    *ineqmax = evopf_ineq(pe_max, pe_min, y, (real*)0, d);")
        e("    const real minpiv = evopf_factor(pt, lu, ipiv);")
        e("    evopf_direction_post(y, pt, lu, ipiv, d, out);")
        e("    return minpiv;")
        e("}")
        e("")

    def generate(self):
        self.gen_header()
        self.gen_eval()
        self.gen_factor()
        self.gen_newton()
        self.gen_refine()
        self.gen_resid()
        self.gen_backward()
        self.gen_vjp()
        self.gen_direction()
        return "\n".join(self.out) + "\n"


def op_stats(txt):
    """Static FP32 operation counts of the generated functions (FMA = 2 flop, mul/add = 1) and the executed-flop
    model  F_exec(n_newton, K) = newton_iter * n_newton + corr_step * K + tail  used by bench.py / DESIGN.md."""
    import json
    import re
    per = {}
    for f in re.split(r"\nRPO_HD ", txt)[1:]:
        name = re.match(r"\w+\s+(\w+)", f).group(1)
        body = f[f.index("{"):]
        small = body.count("SINCOS_SMALL(")
        fma = body.count("FMA(") + 9 * small                      # rpo_sincosf_small: 9 FMA + 2 mul
        mul = len(re.findall(r"[\w\]\)] \* [\w\(-]", body)) + body.count("*=") + 2 * small
        add = len(re.findall(r"[\w\]\)] [+-] [\w\(]", body)) + body.count("+=") + body.count("-=")
        per[name] = dict(fma=fma, mul=mul, add=add, flop=2 * fma + mul + add)
    fl = lambda *names: sum(per[n]["flop"] for n in names)
    newton = fl("evopf_eval_full", "evopf_newton_rhs", "evopf_factor", "evopf_solve", "evopf_newton_update")
    post = fl("evopf_direction_pre", "evopf_direction_fin", "evopf_rhs_t", "evopf_solve_t", "evopf_colsums", "evopf_solve")
    corr = fl("evopf_eval_full", "evopf_eq_resid", "evopf_ineq", "evopf_factor") + post + 2 * 43
    bwd = fl("evopf_eval_full", "evopf_factor", "evopf_rhs_t", "evopf_solve_t", "evopf_colsums", "evopf_backward")
    fb = txt[txt.index("RPO_HD void evopf_flows_f64"):]
    fb = fb[:fb.index("\n}\n")]
    dp_fma = fb.count("fma(") + 14 * 2                                 # rpo_sincos_mixed: 2 fma + 2 mul in float64 per bus
    dp_mul = len(re.findall(r"[\w\]\)] \* [\w\(-]", fb)) + 14 * 2
    per["evopf_flows_f64"] = dict(fma=dp_fma, mul=dp_mul, add=0, flop=2 * dp_fma + dp_mul, precision="float64")
    return dict(per_function=per, newton_iter_flop=newton, corr_step_flop=corr, tail_flop=fl("evopf_tail"),
                refine_pass_f32_flop=fl("evopf_solve", "evopf_newton_update", "evopf_tail_refined") + 22 + 14 * 17, refine_pass_f64_flop=2 * dp_fma + dp_mul,
                backward_flop=bwd, note="general-path sincos not counted (small-angle path taken by the data)")


def main():
    g = Gen()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rpo_b200", "csrc", "evopf_gen.cuh")
    txt = g.generate()
    with open(out, "w") as f:
        f.write(txt)
    import json
    with open(out.replace("evopf_gen.cuh", "evopf_gen_stats.json"), "w") as f:
        json.dump(op_stats(txt), f, indent=1)
    print("order", g.order, "nnzY", g.nnz, "nL", g.nL, "nU", g.nU, "lu_ops", g.lu_ops, "lines", txt.count("\n"))


if __name__ == "__main__":
    main()
