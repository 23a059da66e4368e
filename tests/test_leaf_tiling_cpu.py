"""Host emulation of the index arithmetic of mf_leaf_factor_kernel (csrc/multifrontal.cu): the L-shaped region held in
shared memory (Lp = first np columns, Up = first np rows of the other columns, leading dimensions rounded to 4), the
4 x 4 tiling of the rank-16 update with the per-element ownership rules (L-region only, the next diagonal block belongs
to the look-ahead warp) and the single Schur pass over interleaved row pairs.  Checked: every entry is owned exactly
once, every shared-memory read of a tile stays inside the allocation, and the emulated factorisation equals an
unpivoted partial LU.  (The kernel itself is parity-tested on the GPU through every solve.)"""
import numpy as np
import pytest

NB = 16


def ld4(n):
    return (n + 3) & ~3


def emulate(m, npiv, rng):
    A = rng.standard_normal((m, m)) + m * np.eye(m)
    F = A.copy(order="F")
    nu = m - npiv
    ldL, ldU, ldS = ld4(m), ld4(nu), ld4(m)
    Lp = np.full(ldL * npiv, np.nan)
    Up = np.full(ldU * npiv, np.nan)
    Us = np.full(NB * ldS, np.nan)
    for c in range(npiv):
        Lp[c * ldL:c * ldL + m] = F[:, c]
    for t in range(npiv):
        Up[t * ldU:t * ldU + nu] = F[t, npiv:]

    def cget(i, c):
        return Lp[i + c * ldL] if c < npiv else Up[i * ldU + c - npiv]

    def cset(i, c, v):
        if c < npiv:
            Lp[i + c * ldL] = v
        else:
            Up[i * ldU + c - npiv] = v

    def factor_block(k0, kb, D):
        for j in range(NB):
            for i in range(j + 1, NB):
                D[i, j] /= D[j, j]
                D[i, j + 1:] -= D[i, j] * D[j, j + 1:]
        for a in range(kb):
            for b in range(kb):
                Lp[(k0 + a) + (k0 + b) * ldL] = D[a, b]
        return D

    D = np.eye(NB)
    kb = min(NB, npiv)
    D[:kb, :kb] = F[:kb, :kb]                      # look-ahead warp: first block straight from the front
    D = factor_block(0, kb, D)
    for k0 in range(0, npiv, NB):
        kb, k1 = min(NB, npiv - k0), k0 + min(NB, npiv - k0)
        rd = 1.0 / np.diag(D)
        nrest = m - k1
        for task in range(2 * nrest):               # phase 2, all 16 steps unconditionally (identity padding)
            if task < nrest:
                i = k1 + task
                a = [Lp[i + (k0 + jj) * ldL] if jj < kb else 0.0 for jj in range(NB)]
                for t in range(NB):
                    a[t] *= rd[t]
                    for jj in range(t + 1, NB):
                        a[jj] -= a[t] * D[t, jj]
                for jj in range(kb):
                    Lp[i + (k0 + jj) * ldL] = a[jj]
            else:
                c = k1 + task - nrest
                u = [cget(k0 + t, c) if t < kb else 0.0 for t in range(NB)]
                for t in range(NB):
                    for tt in range(t + 1, NB):
                        u[tt] -= D[tt, t] * u[t]
                for t in range(kb):
                    cset(k0 + t, c, u[t])
                    Us[t * ldS + (c - k1)] = u[t]
        if k1 < npiv:                               # phase 3
            kb2 = min(NB, npiv - k1)
            kend = k1 + kb2
            tr, tcA = (m - k1 + 3) >> 2, (npiv - k1 + 3) >> 2
            trB = tcA
            ntA, ntB = tr * tcA, trB * (tr - tcA)
            owned = set()
            for tile in range(ntA + ntB):
                if tile < ntA:
                    ti, tj = tile % tr, tile // tr
                else:
                    q = tile - ntA
                    ti, tj = q % trB, tcA + q // trB
                i0, c0 = k1 + 4 * ti, k1 + 4 * tj
                assert i0 + 3 < ldL and (c0 - k1) + 3 < ldS          # operand reads stay inside Lp / Us
                if i0 + 3 < kend and c0 + 3 < kend:
                    continue
                for aa in range(4):
                    for b in range(4):
                        i, c = i0 + aa, c0 + b
                        if i < m and c < m and not (i >= npiv and c >= npiv) and not (i < kend and c < kend):
                            assert (i, c) not in owned
                            owned.add((i, c))
                            acc = sum(Lp[i + (k0 + t) * ldL] * Us[t * ldS + (c - k1)] for t in range(NB))
                            cset(i, c, cget(i, c) - acc)
            want = {(i, c) for i in range(k1, m) for c in range(k1, m)
                    if not (i >= npiv and c >= npiv) and not (i < kend and c < kend)}
            assert owned == want
            # look-ahead warp: the next diagonal block, same formula, then factored and published
            D = np.eye(NB)
            for a in range(kb2):
                for b in range(kb2):
                    acc = sum(Lp[(k1 + a) + (k0 + t) * ldL] * Us[t * ldS + b] for t in range(NB))
                    D[a, b] = Lp[(k1 + a) + (k1 + b) * ldL] - acc
            D = factor_block(k1, kb2, D)
    # phase 4: Schur complement, row pairs ti and ti + H counted from the 4-aligned row rb
    rb = npiv & ~3
    npair = (m - rb + 1) >> 1
    H = (npair + 1) >> 1
    tcS = (nu + 3) >> 2
    seen = set()
    for tile in range(H * tcS):
        ti, tj = tile % H, tile // H
        rA, rB, cc0 = rb + 2 * ti, rb + 2 * (ti + H), 4 * tj
        assert rB + 1 < ldL and cc0 + 3 < max(ldU, 4)
        for i in (rA, rA + 1, rB, rB + 1):
            for b in range(4):
                cc = cc0 + b
                if npiv <= i < m and cc < nu:
                    assert (i, cc) not in seen
                    seen.add((i, cc))
                    F[i, npiv + cc] -= sum(Lp[i + t * ldL] * Up[t * ldU + cc] for t in range(npiv))
    assert seen == {(i, cc) for i in range(npiv, m) for cc in range(nu)}
    for c in range(npiv):
        F[:, c] = Lp[c * ldL:c * ldL + m]
    for t in range(npiv):
        F[t, npiv:] = Up[t * ldU:t * ldU + nu]
    R = A.copy()
    for j in range(npiv):
        R[j + 1:, j] /= R[j, j]
        R[j + 1:, j + 1:] -= np.outer(R[j + 1:, j], R[j, j + 1:])
    return np.abs(F - R).max() / np.abs(R).max()


@pytest.mark.parametrize("m,npiv", [(50, 20), (164, 67), (37, 37), (70, 33), (21, 16), (19, 3), (100, 48), (66, 35)])
def test_leaf_kernel_index_logic_reproduces_an_unpivoted_partial_lu(m, npiv):
    assert emulate(m, npiv, np.random.default_rng(m * 1000 + npiv)) < 1e-13


def test_schur_row_pairs_cover_every_row_once_and_stay_inside_the_panel():
    for m in range(17, 200):
        for npiv in range(1, m + 1):
            rb = npiv & ~3
            npair = (m - rb + 1) >> 1
            H = (npair + 1) >> 1
            rows = []
            for ti in range(H):
                for i in (rb + 2 * ti, rb + 2 * ti + 1, rb + 2 * (ti + H), rb + 2 * (ti + H) + 1):
                    assert i < ld4(m)
                    if npiv <= i < m:
                        rows.append(i)
            assert sorted(rows) == list(range(npiv, m))
