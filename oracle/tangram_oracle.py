"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement of the Tangram `Mapper` hot path (reference:
/root/reference/tangram/mapping_optimizer.py, class Mapper, lines 14-408).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; `tangram_b200/` never does.

It is a *closed-form* restatement: the loss, its gradient w.r.t. the mapping matrix M
and the Adam update are written out explicitly (no autograd), so it is an independent
statement of the same algorithm, evaluated in float32 (default) or float64.

Parity pin: `tests/golden/make_golden.py` imports the reference file by path, runs
it on seeded inputs and stores inputs+outputs under `tests/golden/*.npz`;
`tests/test_oracle.py` checks this module against those vectors (and, when
/root/reference is present, against the live reference).  The arithmetic itself lives
in PyTorch (unpinned dependency, setup.py:26; environment.yml pins 1.4.0); semantics
restated here are those of torch 2.11: `cosine_similarity` clamps each norm at 1e-8,
`KLDivLoss(reduction="sum")` is xlogy(t,t) - t*input, `Adam` uses lerp for the first
moment and `denom = sqrt(v)/sqrt(1-b2^t) + eps`, `step = lr/(1-b1^t)`.
"""
import numpy as np
import torch

EPS_COS = 1e-8  # torch.nn.functional.cosine_similarity default eps


def _as_op(mat, dtype, device="cpu"):
    """Dense ndarray / scipy sparse / torch tensor -> (matmul, rmatmul) closures.

    The reference receives dense V x V matrices (mapping_optimizer.py:125-141).
    The oracle also accepts scipy CSR so that large graphs stay cheap.
    """
    if mat is None:
        return None
    if hasattr(mat, "tocsr"):  # scipy sparse
        csr = mat.tocsr()
        t = torch.sparse_csr_tensor(
            torch.as_tensor(csr.indptr, dtype=torch.int64),
            torch.as_tensor(csr.indices, dtype=torch.int64),
            torch.as_tensor(csr.data, dtype=dtype),
            size=csr.shape,
        ).to(device)
        csr_t = csr.T.tocsr()
        tt = torch.sparse_csr_tensor(
            torch.as_tensor(csr_t.indptr, dtype=torch.int64),
            torch.as_tensor(csr_t.indices, dtype=torch.int64),
            torch.as_tensor(csr_t.data, dtype=dtype),
            size=csr_t.shape,
        ).to(device)
        return (lambda x: torch.sparse.mm(t, x)), (lambda x: torch.sparse.mm(tt, x))
    dense = torch.as_tensor(np.asarray(mat), dtype=dtype).to(device)
    return (lambda x: dense @ x), (lambda x: dense.t() @ x)


def _cos_cols(a, b):
    """Column-wise cosine with torch's per-norm clamp; returns (cos[K], na[K], nb[K])."""
    na = torch.clamp(torch.linalg.vector_norm(a, dim=0), min=EPS_COS)
    nb = torch.clamp(torch.linalg.vector_norm(b, dim=0), min=EPS_COS)
    c = (a * b).sum(dim=0) / (na * nb)
    return c, na, nb


def _dcos_cols(a, b, c, na, nb):
    """d(mean_k cos(a_k,b_k))/da  (SURVEY appendix A.2)."""
    n = a.shape[1]
    return (b / (na * nb) - a * (c / (na * na))) / n


class OracleMapper:
    """Closed-form CPU restatement of reference `Mapper` (mapping_optimizer.py:14-408)."""

    def __init__(
        self,
        S,
        G,
        train_genes_idx=None,
        val_genes_idx=None,
        d=None,
        d_source=None,
        lambda_g1=1.0,
        lambda_d=0,
        lambda_g2=0,
        lambda_r=0,
        lambda_l1=0,
        lambda_l2=0,
        lambda_neighborhood_g1=0,
        voxel_weights=None,
        lambda_getis_ord=0,
        lambda_geary=0,
        lambda_moran=0,
        neighborhood_filter=None,
        ct_encode=None,
        lambda_ct_islands=0,
        spatial_weights=None,
        device="cpu",
        adata_map=None,
        random_state=None,
        dtype=torch.float32,
        M0=None,
    ):
        if lambda_geary > 0 or lambda_moran > 0:
            # mapping_optimizer.py:173-185 -- out of scope
            raise NotImplementedError("Moran / Geary terms are outside the hot-path scope")
        self.dtype = dtype
        self.device = device      # "cpu" (default); the GPU tests run this same restatement on "cuda" at BASELINE's full sizes
        self.random_state = random_state

        def _t(x):                # f32 first (the reference's cast, :83-84), then the oracle's working dtype
            x = x.detach() if hasattr(x, "detach") else np.asarray(x)
            return torch.as_tensor(x, dtype=torch.float32).to(device=device, dtype=dtype)
        S, G = _t(S), _t(G)
        # mapping_optimizer.py:87-92: train subset (val subset is never used, :321-322)
        if train_genes_idx is not None:
            S = S[:, train_genes_idx]
            G = G[:, train_genes_idx]
        self.S = S.contiguous()
        self.G = G.contiguous()
        self.lam = dict(
            g1=lambda_g1, d=lambda_d, g2=lambda_g2, r=lambda_r, l1=lambda_l1, l2=lambda_l2,
            nb=lambda_neighborhood_g1, ct=lambda_ct_islands, go=lambda_getis_ord,
        )
        self.d = None if d is None else _t(d)
        self.d_source = None if d_source is None else _t(d_source)
        self.W = _as_op(voxel_weights, dtype, device)
        self.F = _as_op(neighborhood_filter, dtype, device)
        self.A = _as_op(spatial_weights, dtype, device)
        self.E = None if ct_encode is None else _t(ct_encode)

        # constants (mapping_optimizer.py:144, :170-171, :236)
        if self.lam["go"] > 0:
            self.go_ref = self.A[0](self.G) / self.G.sum(dim=0)
        if self.lam["nb"] > 0:
            self.WG = self.W[0](self.G)

        # mapping_optimizer.py:147-157: legacy numpy RNG, float64 draw, cast to f32
        if adata_map is not None:
            raise NotImplementedError
        if M0 is None:
            if self.random_state:
                np.random.seed(seed=self.random_state)
            M0 = np.random.normal(0, 1, (self.S.shape[0], self.G.shape[0]))
        self.M = _t(M0).clone()
        self.m = torch.zeros_like(self.M)
        self.v = torch.zeros_like(self.M)
        self.t = 0

    # ---------------------------------------------------------------- loss + gradient
    def loss_and_grad(self, M=None, need_grad=True):
        """Returns (terms: dict of python floats, dM or None).

        Forward follows mapping_optimizer.py:199-270; backward is appendix A.2 of SURVEY.md.
        """
        lam, S, G = self.lam, self.S, self.G
        M = self.M if M is None else M
        N, V = M.shape
        K = S.shape[1]
        nan = float("nan")

        P = torch.softmax(M, dim=1)                      # :201
        Y = P.t() @ S                                    # :202
        terms = {}

        # gene-voxel cosine (:205, :208)
        c_g, nyg, ngg = _cos_cols(Y, G)
        gv = lam["g1"] * c_g.mean()
        terms["main_loss"] = float(c_g.mean())
        dY = -lam["g1"] * _dcos_cols(Y, G, c_g, nyg, ngg)
        # voxel-gene cosine (:206, :209)
        if lam["g2"] != 0:
            c_v, nyv, ngv = _cos_cols(Y.t(), G.t())
            vg = lam["g2"] * c_v.mean()
            terms["vg_reg"] = float(c_v.mean())
            dY = dY - lam["g2"] * _dcos_cols(Y.t(), G.t(), c_v, nyv, ngv).t()
        else:
            vg = 0.0
            terms["vg_reg"] = nan                         # 0/0 at :209
        total = -gv - vg

        dP_cols = None   # V-vector broadcast down the rows (cells-mode density)
        dP_outer = None  # (N-vector, V-vector) outer product (clusters-mode density)
        # density KL (:212-221)
        if self.d is not None:
            if self.d_source is not None:
                dhat = self.d_source @ P
            else:
                dhat = P.sum(dim=0) / N
            kl = (torch.special.xlogy(self.d, self.d) - self.d * torch.log(dhat)).sum()
            total = total + lam["d"] * kl
            terms["kl_reg"] = float(kl) if lam["d"] != 0 else nan
            gd = -lam["d"] * self.d / dhat                # dL/d dhat_j
            if self.d_source is not None:
                dP_outer = (self.d_source, gd)
            else:
                dP_cols = gd / N
        else:
            terms["kl_reg"] = nan

        # entropy (:224-225)
        if lam["r"] != 0:
            logP = torch.log_softmax(M, dim=1)
            ent = -(logP * P).sum()
            total = total + lam["r"] * ent
            terms["entropy_reg"] = float(ent)
        else:
            terms["entropy_reg"] = nan
        # L1 / L2 (:228-231)
        terms["l1_reg"] = float(M.abs().sum()) if lam["l1"] != 0 else nan
        terms["l2_reg"] = float((M * M).sum()) if lam["l2"] != 0 else nan
        if lam["l1"] != 0:
            total = total + lam["l1"] * M.abs().sum()
        if lam["l2"] != 0:
            total = total + lam["l2"] * (M * M).sum()

        # neighbourhood-weighted cosine (:234-239)
        if lam["nb"] > 0:
            WY = self.W[0](Y)
            c_n, nwy, nwg = _cos_cols(WY, self.WG)
            total = total - lam["nb"] * c_n.mean()
            terms["gv_neighborhood_sim"] = float(c_n.mean())
            dY = dY - lam["nb"] * self.W[1](_dcos_cols(WY, self.WG, c_n, nwy, nwg))
        else:
            terms["gv_neighborhood_sim"] = nan

        # cell-type islands (:242-248)
        dC = None
        if lam["ct"] > 0:
            C = P.t() @ self.E
            R = C - self.F[0](C)
            ct = torch.clamp(R, min=0).mean()
            total = total + lam["ct"] * ct
            terms["ct_island_penalty"] = float(ct)
            H = (R > 0).to(self.dtype) / R.numel()
            dC = lam["ct"] * (H - self.F[1](H))
        else:
            terms["ct_island_penalty"] = nan

        # Getis-Ord G* (:170-171, :251, :255-257)
        if lam["go"] > 0:
            ys = Y.sum(dim=0)
            Q = self.A[0](Y) / ys
            c_q, nq, nr = _cos_cols(Q, self.go_ref)
            total = total - lam["go"] * c_q.mean()
            terms["getis_ord_sim"] = float(c_q.mean())
            gq = _dcos_cols(Q, self.go_ref, c_q, nq, nr)
            dY = dY - lam["go"] * (self.A[1](gq / ys) - ((gq * Q).sum(dim=0) / ys)[None, :])
        else:
            terms["getis_ord_sim"] = nan

        terms["total_loss"] = float(total)
        if not need_grad:
            return terms, None

        # ---- backward to M (appendix A.2)
        dP = S @ dY.t()
        if dP_cols is not None:
            dP = dP + dP_cols[None, :]
        if dP_outer is not None:
            dP = dP + dP_outer[0][:, None] * dP_outer[1][None, :]
        if lam["r"] != 0:
            dP = dP - lam["r"] * (logP + 1.0)
        if dC is not None:
            dP = dP + self.E @ dC.t()
        rowdot = (P * dP).sum(dim=1, keepdim=True)
        dM = P * (dP - rowdot)
        if lam["l1"] != 0:
            dM = dM + lam["l1"] * torch.sign(M)
        if lam["l2"] != 0:
            dM = dM + 2.0 * lam["l2"] * M
        return terms, dM

    # ---------------------------------------------------------------- Adam
    def adam_step(self, g, lr, b1=0.9, b2=0.999, eps=1e-8):
        """torch.optim.Adam single-tensor path (torch/optim/adam.py, torch 2.11), restated."""
        self.t += 1
        t = self.t
        self.m = self.m + (g - self.m) * (1 - b1)                 # lerp_
        self.v = self.v * b2 + (1 - b2) * g * g                   # mul_ + addcmul_
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        step_size = lr / bc1
        denom = self.v.sqrt() / (bc2 ** 0.5) + eps
        self.M = self.M - step_size * (self.m / denom)            # addcdiv_

    # ---------------------------------------------------------------- train loop
    def train(self, num_epochs, learning_rate=0.1, print_each=100, val_each=None):
        """mapping_optimizer.py:358-408.  Returns (softmax(M) as f32 ndarray, history)."""
        keys = ["total_loss", "main_loss", "vg_reg", "kl_reg", "entropy_reg"]
        val_keys = ["val_total_loss", "val_gene_sim", "val_sp_sparsity_weighted_sim", "val_entropy"]
        history = {k: [] for k in keys + val_keys}
        self.terms_history = []
        # :373 -- `torch.optim.Adam([self.M], lr)` is built inside train(): every call starts from zero moments, t = 0
        self.m = torch.zeros_like(self.M)
        self.v = torch.zeros_like(self.M)
        self.t = 0
        for t in range(num_epochs):
            terms, dM = self.loss_and_grad()
            self.terms_history.append(terms)
            history["total_loss"].append(np.array(terms["total_loss"], dtype=np.float32))
            for k in keys[1:]:
                history[k].append(terms[k])
            if print_each is not None and t % print_each == 0:
                print(format_terms(terms))
            self.adam_step(dM, learning_rate)
            if val_each is not None and t % val_each == 0:
                vals = self.val_terms()
                for k, x in zip(val_keys, vals):
                    history[k].append(x)
        out = torch.softmax(self.M, dim=1).to(torch.float32).cpu().numpy()
        return out, history

    def val_terms(self):
        """mapping_optimizer.py:311-356 (uses the *train* matrices, :321-322)."""
        P = torch.softmax(self.M, dim=1)
        Y = P.t() @ self.S
        G = self.G
        c_g, _, _ = _cos_cols(Y, G)
        c_v, _, _ = _cos_cols(Y.t(), G.t())
        gv_sim = float(c_g.mean())
        vg_sim = float(c_v.mean())
        w = (G != 0).to(self.dtype).sum(dim=0) / G.shape[0]      # 1 - gene_sparsity
        sp = float((c_g * w / w.sum()).sum())
        ent = float(-((torch.log(P) * P).sum(dim=1) / np.log(P.shape[1])).mean())
        return gv_sim + vg_sim, gv_sim, sp, ent


_TERM_NAMES = [
    ("main_loss", "Gene-voxel score"),
    ("vg_reg", "Voxel-gene score"),
    ("kl_reg", "Cell densities reg"),
    ("entropy_reg", "Entropy reg"),
    ("l1_reg", "L1 reg"),
    ("l2_reg", "L2 reg"),
    ("gv_neighborhood_sim", "Spatial weighted score"),
    ("ct_island_penalty", "Cell type islands penalty"),
    ("getis_ord_sim", "Getis-Ord score"),
]


def format_terms(terms):
    """The reference's per-epoch print line (mapping_optimizer.py:272-307), NaN terms dropped."""
    msg = ["{}: {:.3f}".format(name, terms[k]) for k, name in _TERM_NAMES if not np.isnan(terms[k])]
    return str(msg).replace("[", "").replace("]", "").replace("'", "")


# -------------------------------------------------------------------- synthetic inputs
def synthetic_inputs(n_cells, n_voxels, n_genes, seed=0, n_types=0, clusters=False):
    """Seeded expression-like inputs, SURVEY.md 8(d).  Shared by tests and bench."""
    rng = np.random.default_rng(seed)
    S = np.log1p(rng.poisson(0.6, (n_cells, n_genes))).astype(np.float32)
    G = np.log1p(rng.poisson(2.0, (n_voxels, n_genes))).astype(np.float32)
    S[:, ~S.any(axis=0)] = 1.0
    G[:, ~G.any(axis=0)] = 1.0
    out = dict(S=S, G=G)
    if clusters:
        w = rng.random(n_cells) + 0.1
        out["d_source"] = (w / w.sum()).astype(np.float32)
        out["d"] = (np.ones(n_voxels) / n_voxels).astype(np.float32)
    else:
        out["d"] = (G.sum(axis=1) / G.sum()).astype(np.float32)
    if n_types:
        lab = rng.integers(0, n_types, n_cells)
        E = np.zeros((n_cells, n_types), dtype=np.float32)
        E[np.arange(n_cells), lab] = 1.0
        out["ct_encode"] = E
    return out


def grid_graph(n_voxels):
    """Voxels on a ceil(sqrt(V)) square grid; 4-neighbour graph as scipy CSR pair
    (connectivities, distances), the shape `squidpy.gr.spatial_neighbors` leaves in
    adata_sp.obsp (mapping_utils.py:95-100)."""
    import scipy.sparse as sp
    side = int(np.ceil(np.sqrt(n_voxels)))
    idx = np.arange(n_voxels)
    x, y = idx % side, idx // side
    rows, cols, dist = [], [], []
    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1)):
        nx, ny = x + dx, y + dy
        j = ny * side + nx
        ok = (nx >= 0) & (nx < side) & (ny >= 0) & (j < n_voxels) & (j >= 0)
        rows.append(idx[ok]); cols.append(j[ok])
        dist.append(np.full(ok.sum(), np.hypot(dx, dy)))
    rows, cols, dist = map(np.concatenate, (rows, cols, dist))
    conn = sp.csr_matrix((np.ones_like(dist), (rows, cols)), shape=(n_voxels, n_voxels))
    dmat = sp.csr_matrix((dist, (rows, cols)), shape=(n_voxels, n_voxels))
    return conn, dmat


def spatial_weights_from_graph(conn, dist, standardized, self_inclusion):
    """Restates tangram/spatial_weights.py:5-30 on scipy CSR (no libpysal/sklearn):
    standardized -> row-L1-normalised distances on the connectivity pattern;
    else binary connectivities; optional + I.  Returns scipy CSR (float32)."""
    import scipy.sparse as sp
    if standardized:
        g = dist.tocsr().astype(np.float64).copy()
        rs = np.asarray(abs(g).sum(axis=1)).ravel()
        rs[rs == 0] = 1.0
        g = sp.diags(1.0 / rs) @ g
        w = g.multiply(conn.tocsr() != 0).tocsr()
    else:
        w = conn.tocsr().astype(np.float64).copy()
    if self_inclusion:
        w = w + sp.identity(w.shape[0], format="csr")
    return w.tocsr().astype(np.float32)


class OracleMapperConstrained:
    """Closed-form CPU restatement of reference `MapperConstrained` (mapping_optimizer.py:411-639): the mapping
    matrix M plus a per-cell filter F (sigmoid), Adam over [M, F].  TEST INFRASTRUCTURE ONLY.

    Quirks preserved: M is drawn twice and the second draw is used (:475, :485); F is drawn after M (:490);
    `entropy_reg` is logged as +sum(P log P) (:526, :540); history values are strings (:630)."""

    def __init__(self, S, G, d, lambda_d=1, lambda_g1=1, lambda_g2=1, lambda_r=0, lambda_count=1, lambda_f_reg=1,
                 target_count=None, device="cpu", adata_map=None, random_state=None, dtype=torch.float32,
                 M0=None, F0=None):
        if adata_map is not None:
            raise NotImplementedError
        self.dtype = dtype
        self.S = torch.as_tensor(np.asarray(S), dtype=torch.float32).to(dtype)
        self.G = torch.as_tensor(np.asarray(G), dtype=torch.float32).to(dtype)
        self.d = None if d is None else torch.as_tensor(np.asarray(d), dtype=torch.float32).to(dtype)
        self.lam = dict(d=lambda_d, g1=lambda_g1, g2=lambda_g2, r=lambda_r, c=lambda_count, f=lambda_f_reg)
        self.target_count = self.G.shape[0] if target_count is None else target_count
        N, V = self.S.shape[0], self.G.shape[0]
        if M0 is None or F0 is None:
            if random_state:
                np.random.seed(seed=random_state)
            np.random.normal(0, 1, (N, V))          # first draw, discarded by the reference (:475 then :485)
            M0 = np.random.normal(0, 1, (N, V))
            F0 = np.random.normal(0, 1, N)
        self.M = torch.as_tensor(np.asarray(M0), dtype=torch.float32).to(dtype).clone()
        self.F = torch.as_tensor(np.asarray(F0), dtype=torch.float32).to(dtype).clone()
        self.mM, self.vM = torch.zeros_like(self.M), torch.zeros_like(self.M)
        self.mF, self.vF = torch.zeros_like(self.F), torch.zeros_like(self.F)
        self.t = 0

    def loss_and_grad(self):
        lam, S, G = self.lam, self.S, self.G
        N, V = self.M.shape
        K = S.shape[1]
        P = torch.softmax(self.M, dim=1)                       # :506
        f = torch.sigmoid(self.F)                              # :507
        Sf = S * f[:, None]                                    # :519
        Y = P.t() @ Sf                                         # :521
        terms = {}
        c_g, nyg, ngg = _cos_cols(Y, G)
        c_v, nyv, ngv = _cos_cols(Y.t(), G.t())
        gv, vg = lam["g1"] * c_g.mean(), lam["g2"] * c_v.mean()
        terms["main_loss"] = float(c_g.mean())
        terms["vg_reg"] = float(c_v.mean()) if lam["g2"] != 0 else float("nan")
        dY = -lam["g1"] * _dcos_cols(Y, G, c_g, nyg, ngg)
        if lam["g2"] != 0:
            dY = dY - lam["g2"] * _dcos_cols(Y.t(), G.t(), c_v, nyv, ngv).t()
        total = -gv - vg
        s = f.sum()
        df_extra = torch.zeros_like(f)
        dP_cols = None
        if self.d is not None:                                 # :511-515
            csf = (P * f[:, None]).sum(dim=0)
            dhat = csf / s
            kl = (torch.special.xlogy(self.d, self.d) - self.d * torch.log(dhat)).sum()
            total = total + lam["d"] * kl
            terms["kl_reg"] = float(kl) if lam["d"] != 0 else float("nan")
            g_cs = -lam["d"] * self.d / csf                    # dL / d csf_j
            dP_cols = g_cs                                     # times f_i below
            df_extra = df_extra + (P @ g_cs) + lam["d"] * self.d.sum() / s
        else:
            terms["kl_reg"] = float("nan")
        plogp = (torch.log_softmax(self.M, dim=1) * P).sum()
        total = total - lam["r"] * plogp                       # :526, :575
        terms["entropy_reg"] = float(plogp) if lam["r"] != 0 else float("nan")
        cnt = s - self.target_count                            # :528-529
        total = total + lam["c"] * cnt.abs()
        terms["count_reg"] = float(cnt.abs()) if lam["c"] != 0 else float("nan")
        freg = (f - f * f).sum()                               # :531-532
        total = total + lam["f"] * freg
        terms["lambda_f_reg"] = float(freg) if lam["f"] != 0 else float("nan")
        terms["total_loss"] = float(total)
        # backward
        SdY = S @ dY.t()                                       # (N, V): d Y-terms / d (f_i P_ij)
        dP = f[:, None] * SdY
        if dP_cols is not None:
            dP = dP + f[:, None] * dP_cols[None, :]
        if lam["r"] != 0:
            dP = dP - lam["r"] * (torch.log_softmax(self.M, dim=1) + 1.0)
        dM = P * (dP - (P * dP).sum(dim=1, keepdim=True))
        df = (P * SdY).sum(dim=1) + df_extra + lam["c"] * torch.sign(cnt) + lam["f"] * (1 - 2 * f)
        dF = df * f * (1 - f)
        return terms, dM, dF

    def _adam(self, x, g, m, v, lr, b1=0.9, b2=0.999, eps=1e-8):
        m = m + (g - m) * (1 - b1)
        v = v * b2 + (1 - b2) * g * g
        step = lr / (1 - b1 ** self.t)
        x = x - step * (m / (v.sqrt() / ((1 - b2 ** self.t) ** 0.5) + eps))
        return x, m, v

    def train(self, num_epochs, learning_rate=0.1, print_each=100):
        keys = ["total_loss", "main_loss", "vg_reg", "kl_reg", "entropy_reg", "count_reg", "lambda_f_reg"]
        hist = {k: [] for k in keys}
        self.float_history = {k: [] for k in keys}
        for t in range(num_epochs):
            terms, dM, dF = self.loss_and_grad()
            for k in keys:
                self.float_history[k].append(terms[k])
                hist[k].append(format_constrained_value(k, terms[k]))
            if print_each is not None and t % print_each == 0:
                print(format_constrained_terms(terms))
            self.t += 1
            self.M, self.mM, self.vM = self._adam(self.M, dM, self.mM, self.vM, learning_rate)
            self.F, self.mF, self.vF = self._adam(self.F, dF, self.mF, self.vF, learning_rate)
        out = torch.softmax(self.M, dim=1).to(torch.float32).numpy()
        return out, torch.sigmoid(self.F).to(torch.float32).numpy(), hist


def format_constrained_value(key, x):
    """History entries of MapperConstrained are `str(...)` of what _loss_fn returns (:630): a tensor repr for
    total_loss, python floats (or nan) for the rest."""
    if key == "total_loss":
        return "tensor({:.4f}, grad_fn=<AddBackward0>)".format(x)
    return str(x)


def format_constrained_terms(terms):
    """The print line of MapperConstrained (:546-573)."""
    names = [("main_loss", "Score"), ("vg_reg", "VG reg"), ("kl_reg", "KL reg"), ("entropy_reg", "Entropy reg"),
             ("count_reg", "Count reg"), ("lambda_f_reg", "Lambda f reg")]
    msg = ["{}: {:.3f}".format(n, terms[k]) for k, n in names if not np.isnan(terms[k])]
    return str(msg).replace("[", "").replace("]", "").replace("'", "")
