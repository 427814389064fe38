"""TEST INFRASTRUCTURE / DESIGN VALIDATION -- the O(1)-per-observation ("scan") form of the B0 (cell-integrated
Matern-1/2) family, stated in torch and checked against the dense features of the reference
(gridded_kronecker_structure.py:1325-1374, restated in vggp_oracle.b0_features_dense).  Not on any product path: it
pins the algebra the next per-observation kernel of this family will implement.

For x with containing cell c (c = -1 left of the mesh, c = K-1 right of it; cells i = 0..K-2, knots t_0..t_{K-1}):
    phi_i(x) = fL(x) * GL[c][i]   (i < c),    fC(x)   (i == c),    fR(x) * GR[c][i]   (i > c)
    fL = s2 l exp(-(x - t_c) / l),   fR = s2 l exp(-(t_{c+1} - x) / l),   fC = 2 s2 l - fL - fR
    GL[c][i] = exp(-(t_c - t_{i+1}) / l) - exp(-(t_c - t_i) / l)          (the cells to the left decay from t_c)
    GR[c][i] = exp(-(t_i - t_{c+1}) / l) - exp(-(t_{i+1} - t_{c+1}) / l)  (the cells to the right decay from t_{c+1})
so every contraction of phi with a grid-side tensor is three per-cell table entries times (fL, fC, fR):
    mu   = sum_{X,Y} f1^X f2^Y T^{XY}[c1][c2],     T^{XY} = G1^X A G2^Y^T                  (9 tables of (M1+2)(M2+2))
    p_d  = sum_{X,Y} f^X f^Y W^{XY}[c],            W^{XY}[c] = (G^X P_d G^Y^T)[c][c]        (6 distinct per matrix)
with G^C[c] = e_c (the unit row).  Per observation: 2 exponentials and ~45 multiply-adds per dimension, independent of
the grid size; the tables cost O(M sum_d M_d) per step on the grid side (dense GEMMs, tensor cores)."""
import torch


def transform_matrices(mesh: torch.Tensor, l: torch.Tensor):
    """GL, GC, GR: (K+1, K-1) each; row e = c + 1 for the containing cell c = -1..K-1."""
    t = mesh.to(torch.float64)
    K = t.numel()
    M = K - 1
    c = torch.arange(-1, K).reshape(-1, 1)                  # (K+1, 1)
    i = torch.arange(M).reshape(1, -1)                      # (1, M)
    tc = t[c.clamp(0, K - 1)]                               # lower knot of the cell (t_{K-1} right of the mesh)
    tc1 = t[(c + 1).clamp(0, K - 1)]                        # upper knot of the cell (t_0 left of the mesh)
    ti, ti1 = t[i], t[i + 1]
    left = i < c
    right = i > c
    GL = torch.where(left, torch.exp(-(tc - ti1).clamp(min=0) / l) - torch.exp(-(tc - ti).clamp(min=0) / l), torch.zeros(()))
    GR = torch.where(right, torch.exp(-(ti - tc1).clamp(min=0) / l) - torch.exp(-(ti1 - tc1).clamp(min=0) / l), torch.zeros(()))
    GC = (i == c).to(torch.float64)
    return GL, GC, GR


def local_features(mesh: torch.Tensor, x: torch.Tensor, l: torch.Tensor, s2: torch.Tensor):
    """e = c + 1 in [0, K] and (fL, fC, fR) per observation; entries that have no cells on their side are zero."""
    t = mesh.to(x.dtype)
    K = t.numel()
    idx = torch.searchsorted(t, x, right=False)             # knots < x   ->  x in (t_{idx-1}, t_idx]
    c = idx - 1                                             # -1 .. K-1
    tc = t[c.clamp(0, K - 1)]
    tc1 = t[(c + 1).clamp(0, K - 1)]
    fL = s2 * l * torch.exp(-(x - tc) / l)
    fR = s2 * l * torch.exp(-(tc1 - x) / l)
    real = (c >= 0) & (c <= K - 2)
    fC = torch.where(real, 2 * s2 * l - fL - fR, torch.zeros((), dtype=x.dtype))
    fL = torch.where(c >= 0, fL, torch.zeros((), dtype=x.dtype))
    fR = torch.where(c <= K - 2, fR, torch.zeros((), dtype=x.dtype))
    return c + 1, torch.stack([fL, fC, fR])                 # (N,), (3, N)


def features_from_scan(mesh, x, l, s2):
    """Dense (K-1, N) features rebuilt from the scan form (only to compare with the reference's)."""
    e, f = local_features(mesh, x, l, s2)
    GL, GC, GR = transform_matrices(mesh, l)
    return (f[0] * GL[e].T + f[1] * GC[e].T + f[2] * GR[e].T)


def mean_p_q_scan(meshes, X, l, s2, A, Ps, Qs):
    """mu, prod_d p_d, prod_d q_d per observation through the per-cell tables (D = 1 or 2)."""
    D = len(meshes)
    es, fs, Gs = [], [], []
    for d in range(D):
        e, f = local_features(meshes[d], X[:, d], l[d], s2[d])
        es.append(e)
        fs.append(f)
        Gs.append(transform_matrices(meshes[d], l[d]))
    if D == 1:
        T = torch.stack([G @ A for G in Gs[0]])                           # (3, E1)
        mu = (fs[0] * T[:, es[0]]).sum(0)
    else:
        T = torch.stack([torch.stack([Gx @ A @ Gy.T for Gy in Gs[1]]) for Gx in Gs[0]])   # (3, 3, E1, E2)
        mu = torch.einsum("xn,yn,xyn->n", fs[0], fs[1], T[:, :, es[0], es[1]])
    p = torch.ones_like(mu)
    q = torch.ones_like(mu)
    for d in range(D):
        for mat, acc in ((Ps[d], "p"), (Qs[d], "q")):
            W = torch.stack([torch.stack([((Gx @ mat) * Gy).sum(1) for Gy in Gs[d]]) for Gx in Gs[d]])   # (3, 3, E)
            v = torch.einsum("xn,yn,xyn->n", fs[d], fs[d], W[:, :, es[d]])
            if acc == "p":
                p = p * v
            else:
                q = q * v
    return mu, p, q
