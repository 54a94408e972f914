"""Independent dense assembly of the SIPG Laplace matrix (numpy, tiny meshes only).

TEST INFRASTRUCTURE.  Derived directly from the bilinear form

    a(u,v) = sum_K (grad v, grad u)_K - sum_f <{grad v}.n [u]> - sum_f <[v] {grad u}.n> + sum_f tau_f <[v][u]>

with [u] = u^- - u^+, {.} the mean, n = n^-; Dirichlet faces: [u] = 2u^-... i.e. the
mirror principle gives -<dn v, u> - <v, dn u> + 2 tau <v,u>; Neumann faces: nothing.
It shares no code with oracle/sipg_oracle.c: 1-D tables come from numpy.polynomial,
shape functions are evaluated as full Kronecker products (no sum factorisation) and
faces are visited once per unique face.  Used by tests/test_oracle_dense.py to check the
matrix-free oracle column by column.
"""
import numpy as np
from numpy.polynomial import legendre as npleg


def gauss01(n):
    x, w = npleg.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def lobatto01(n):
    if n == 2:
        return np.array([0.0, 1.0])
    c = np.zeros(n)
    c[n - 1] = 1.0  # P_{n-1}
    interior = np.sort(npleg.legroots(npleg.legder(c)))
    return 0.5 * (np.concatenate([[-1.0], interior, [1.0]]) + 1.0)


def lagrange_tables(nodes, x):
    """V[q,j] = l_j(x_q), G[q,j] = l_j'(x_q) via barycentric-free explicit products."""
    n = len(nodes)
    x = np.atleast_1d(x)
    V = np.ones((len(x), n))
    G = np.zeros((len(x), n))
    for j in range(n):
        for i in range(n):
            if i != j:
                V[:, j] *= (x - nodes[i]) / (nodes[j] - nodes[i])
        for m in range(n):
            if m == j:
                continue
            t = np.full(len(x), 1.0 / (nodes[j] - nodes[m]))
            for i in range(n):
                if i != j and i != m:
                    t *= (x - nodes[i]) / (nodes[j] - nodes[i])
            G[:, j] += t
    return V, G


def kron3(az, ay, ax):
    """rows = points (x fastest), cols = basis (x fastest)"""
    return np.kron(az, np.kron(ay, ax))


def assemble(degree, xmap, nb, nbface, bt, mapping_degree, ip_factor=1.0):
    """Dense SIPG matrix for the mesh arrays (as returned by OracleOperator.mesh())."""
    n = degree + 1
    n3 = n ** 3
    nc = xmap.shape[0]
    N = nc * n3
    xn = lobatto01(n) if degree > 0 else np.array([0.5])
    xq, wq = gauss01(n)
    gl = lobatto01(mapping_degree + 1)
    S, D = lagrange_tables(xn, xq)
    MS, MD = lagrange_tables(gl, xq)
    # cell tables
    PHI_G = [kron3(S, S, D), kron3(S, D, S), kron3(D, S, S)]  # d/dxi_e of phi at cell q-points
    MAP_G = [kron3(MS, MS, MD), kron3(MS, MD, MS), kron3(MD, MS, MS)]
    W3 = np.kron(wq, np.kron(wq, wq))
    A = np.zeros((N, N))

    def face_tables(nodes, f):
        """value and 3 reference-gradient tables of the tensor basis on `nodes` at the face q-points of face f"""
        d, s = f // 2, f % 2
        Vs, Gs = lagrange_tables(nodes, np.array([float(s)]))
        Vq, Gq = lagrange_tables(nodes, xq)
        val1 = [Vq, Vq, Vq]
        der1 = [Gq, Gq, Gq]
        val1[d], der1[d] = Vs, Gs
        # point ordering on the face: lower tangential direction fastest -> kron with singleton in d keeps that
        val = kron3(val1[2], val1[1], val1[0])
        grads = []
        for e in range(3):
            t = list(val1)
            t[e] = der1[e]
            grads.append(kron3(t[2], t[1], t[0]))
        return val, grads

    # geometry helpers
    def cell_jac(c, tabs):
        X = xmap[c]  # (np3,3)
        J = np.stack([tabs[e] @ X for e in range(3)], axis=2)  # (pts, i, e) = dx_i/dxi_e
        return J

    tau = np.zeros(nc)
    face_geo = {}
    w2 = np.kron(wq, wq)
    for c in range(nc):
        J = cell_jac(c, MAP_G)
        det = np.linalg.det(J)
        vol = np.sum(det * W3)
        Jinv = np.linalg.inv(J)  # (pts, e, i) = dxi_e/dx_i
        # physical gradients of all basis functions: (pts, i, j) = sum_e Jinv[e,i] dphi_j/dxi_e
        gphys = sum(Jinv[:, e, :, None] * PHI_G[e][:, None, :] for e in range(3))
        blk = np.einsum("qij,qik,q->jk", gphys, gphys, det * W3)
        A[c * n3:(c + 1) * n3, c * n3:(c + 1) * n3] += blk
        surf = 0.0
        for f in range(6):
            d, s = f // 2, f % 2
            _, mg = face_tables(gl, f)
            Jf = np.stack([mg[e] @ xmap[c] for e in range(3)], axis=2)
            detf = np.linalg.det(Jf)
            Jfi = np.linalg.inv(Jf)
            nv = Jfi[:, d, :] * (1.0 if s else -1.0)  # J^{-T} n_ref
            ln = np.linalg.norm(nv, axis=1)
            nrm = nv / ln[:, None]
            jxw = np.abs(detf) * ln * w2
            pv, pg = face_tables(xn, f)
            gp = sum(Jfi[:, e, :, None] * pg[e][:, None, :] for e in range(3))  # (pts, i, j)
            face_geo[(c, f)] = (nrm, jxw, pv, gp)
            surf += np.sum(jxw) * (1.0 if bt[c, f] != 0 else 0.5)
        tau[c] = surf / vol

    pen = ip_factor * (degree + 1.0) ** 2
    for c in range(nc):
        for f in range(6):
            p = int(nb[c, f])
            sl_m = slice(c * n3, (c + 1) * n3)
            nrm, jxw, vm, gm = face_geo[(c, f)]
            dn_m = np.einsum("qij,qi->qj", gm, nrm)
            if p < 0:
                if bt[c, f] == 1:  # Dirichlet
                    tf = tau[c] * pen
                    A[sl_m, sl_m] += -(dn_m.T * jxw) @ vm - (vm.T * jxw) @ dn_m + 2.0 * tf * (vm.T * jxw) @ vm
                continue
            fp = int(nbface[c, f])
            if not (c < p or (c == p and f > fp)):
                continue
            sl_p = slice(p * n3, (p + 1) * n3)
            _, _, vp, gp = face_geo[(p, fp)]
            dn_p = np.einsum("qij,qi->qj", gp, nrm)  # derivative along n^-
            tf = max(tau[c], tau[p]) * pen
            # jump [u] = Jm u^- + Jp u^+ with Jm = vm, Jp = -vp ; mean normal gradient = 0.5(dn_m u^- + dn_p u^+)
            sides = [(sl_m, vm, 0.5 * dn_m), (sl_p, -vp, 0.5 * dn_p)]
            for (ri, jv_i, av_i) in sides:  # test function v
                for (ci, jv_j, av_j) in sides:  # trial function u
                    A[ri, ci] += -(av_i.T * jxw) @ jv_j - (jv_i.T * jxw) @ av_j + tf * (jv_i.T * jxw) @ jv_j
    return A, tau
