// See multifrontal.hpp: host symbolic analysis + host restatement of the numeric phase.
#include "multifrontal.hpp"

#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <numeric>

namespace ocp {

namespace {

struct Builder {
    int n;
    const int *rowptr, *col;
    const double *xy;
    const uint8_t *kind;
    int leaf;
    std::vector<int> side;
    // tree in postorder
    std::vector<std::vector<int>> node_cols;
    std::vector<std::pair<int, int>> node_children;   // -1 when absent

    void order_block(std::vector<int> &set, int ax) {
        std::stable_sort(set.begin(), set.end(), [&](int a, int b) {
            if (kind[a] != kind[b]) return kind[a] < kind[b];
            if (ax >= 0 && xy[2 * a + ax] != xy[2 * b + ax]) return xy[2 * a + ax] < xy[2 * b + ax];
            return a < b;
        });
    }

    int emit(std::vector<int> &cols, int c0, int c1, int ax) {
        order_block(cols, ax);
        node_cols.push_back(cols);
        node_children.push_back({c0, c1});
        return (int)node_cols.size() - 1;
    }

    int rec(std::vector<int> &set) {
        if ((int)set.size() <= leaf) return emit(set, -1, -1, -1);
        double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
        for (int i : set)
            for (int d = 0; d < 2; ++d) {
                lo[d] = std::min(lo[d], xy[2 * i + d]);
                hi[d] = std::max(hi[d], xy[2 * i + d]);
            }
        const int ax = (hi[0] - lo[0] >= hi[1] - lo[1]) ? 0 : 1;
        std::vector<double> c(set.size());
        for (size_t k = 0; k < set.size(); ++k) c[k] = xy[2 * set[k] + ax];
        std::nth_element(c.begin(), c.begin() + c.size() / 2, c.end());
        const double med = c[c.size() / 2];
        std::vector<int> left, right, sep, rest;
        for (int i : set) {
            if (xy[2 * i + ax] < med) {
                side[i] = 1;
                left.push_back(i);
            } else {
                side[i] = 2;
                right.push_back(i);
            }
        }
        if (left.empty() || right.empty()) {
            for (int i : set) side[i] = 0;
            return emit(set, -1, -1, -1);
        }
        for (int i : right) {
            bool touches = false;
            for (int p = rowptr[i]; p < rowptr[i + 1] && !touches; ++p) touches = (side[col[p]] == 1);
            (touches ? sep : rest).push_back(i);
        }
        for (int i : set) side[i] = 0;
        set.clear();
        set.shrink_to_fit();
        const int c0 = rec(left);
        const int c1 = rest.empty() ? -1 : rec(rest);
        return emit(sep, c0, c1, 1 - ax);
    }
};

}  // namespace

int mf_leaf_size() {
    // 80: on the 32 x 32 reference mesh 128 leaves (one wave of CTAs) and 8 tree levels instead of 158 / 9 with 48;
    // measured step 3.15 ms vs 3.20 ms (profiles/README.md, round 2: leaf-size sweep 48 ... 128)
    if (const char *e = getenv("OCP_MF_LEAF")) return std::max(8, atoi(e));
    return 80;
}

void mf_analyse(int n, const int *rowptr, const int *col, const double *xy, const uint8_t *kind, int leaf,
                MFSymbolic &S) {
    Builder b{n, rowptr, col, xy, kind, leaf, std::vector<int>(n, 0), {}, {}};
    std::vector<int> all(n);
    std::iota(all.begin(), all.end(), 0);
    b.rec(all);
    const int nn = (int)b.node_cols.size();
    // ---- lift every pressure-like dof to the highest tree node among its structural neighbours, so that all
    // velocities it couples to are eliminated before it and its pivot is a (non-zero) Schur complement even
    // with pivoting restricted to the front's own rows.  The neighbours' nodes lie on one root path (they are
    // all adjacent to the dof itself), so "highest" is well defined and the tree stays a valid elimination tree.
    {
        std::vector<int> par(nn, -1), depth(nn, 0), node_of(n, -1);
        for (int s = 0; s < nn; ++s)
            for (int c : {b.node_children[s].first, b.node_children[s].second})
                if (c >= 0) par[c] = s;
        for (int s = nn - 1; s >= 0; --s) depth[s] = par[s] < 0 ? 0 : depth[par[s]] + 1;   // parents have larger ids
        for (int s = 0; s < nn; ++s)
            for (int d : b.node_cols[s]) node_of[d] = s;
        std::vector<int> target(n, -1);
        bool any = false;
        for (int d = 0; d < n; ++d) {
            if (!kind[d]) continue;
            int best = node_of[d];
            for (int p = rowptr[d]; p < rowptr[d + 1]; ++p) {
                const int t = node_of[col[p]];
                if (depth[t] < depth[best]) best = t;
            }
            if (best != node_of[d]) {
                target[d] = best;
                any = true;
            }
        }
        if (any) {
            for (int s = 0; s < nn; ++s) {
                std::vector<int> keep;
                for (int d : b.node_cols[s])
                    if (target[d] < 0) keep.push_back(d);
                b.node_cols[s].swap(keep);
            }
            for (int d = 0; d < n; ++d)
                if (target[d] >= 0) b.node_cols[target[d]].push_back(d);
            for (int s = 0; s < nn; ++s)
                std::stable_sort(b.node_cols[s].begin(), b.node_cols[s].end(),
                                 [&](int a, int c) { return kind[a] < kind[c]; });
        }
    }
    S = MFSymbolic();
    S.n = n;
    S.nnodes = nn;
    S.perm.reserve(n);
    S.first.resize(nn);
    S.np.resize(nn);
    S.m.resize(nn);
    S.parent.assign(nn, -1);
    std::vector<int> pos(n), node_of_pos(n);
    for (int s = 0; s < nn; ++s) {
        S.first[s] = (int)S.perm.size();
        S.np[s] = (int)b.node_cols[s].size();
        for (int d : b.node_cols[s]) {
            pos[d] = (int)S.perm.size();
            node_of_pos[pos[d]] = s;
            S.perm.push_back(d);
        }
        for (int c : {b.node_children[s].first, b.node_children[s].second})
            if (c >= 0) S.parent[c] = s;
    }
    // children lists
    S.child_ptr.assign(nn + 1, 0);
    for (int s = 0; s < nn; ++s)
        if (S.parent[s] >= 0) S.child_ptr[S.parent[s] + 1]++;
    for (int s = 0; s < nn; ++s) S.child_ptr[s + 1] += S.child_ptr[s];
    S.child.resize(S.child_ptr[nn]);
    {
        std::vector<int> next(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (int s = 0; s < nn; ++s)
            if (S.parent[s] >= 0) S.child[next[S.parent[s]]++] = s;
    }
    // update sets (as elimination positions, ascending) in postorder
    std::vector<std::vector<int>> upd(nn);
    std::vector<int> mark(n, -1);
    for (int s = 0; s < nn; ++s) {
        const int last = S.first[s] + S.np[s];
        std::vector<int> &u = upd[s];
        for (int k = S.first[s]; k < last; ++k) {
            const int d = S.perm[k];
            for (int p = rowptr[d]; p < rowptr[d + 1]; ++p) {
                const int q = pos[col[p]];
                if (q >= last && mark[q] != s) {
                    mark[q] = s;
                    u.push_back(q);
                }
            }
        }
        for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci)
            for (int q : upd[S.child[ci]])
                if (q >= last && mark[q] != s) {
                    mark[q] = s;
                    u.push_back(q);
                }
        std::sort(u.begin(), u.end());
        S.m[s] = S.np[s] + (int)u.size();
    }
    // front index lists and offsets
    S.idx_ptr.assign(nn + 1, 0);
    S.front_ptr.assign(nn + 1, 0);
    for (int s = 0; s < nn; ++s) {
        S.idx_ptr[s + 1] = S.idx_ptr[s] + S.m[s];
        S.front_ptr[s + 1] = S.front_ptr[s] + (long long)S.m[s] * S.m[s];
        S.max_front = std::max(S.max_front, S.m[s]);
        S.max_np = std::max(S.max_np, S.np[s]);
        const double p = S.np[s], mm = S.m[s];
        S.flops += 2.0 * p * mm * mm - 2.0 * p * p * mm + 2.0 / 3.0 * p * p * p;
    }
    S.fsize = S.front_ptr[nn];
    S.idx.resize(S.idx_ptr[nn]);
    for (int s = 0; s < nn; ++s) {
        int *I = S.idx.data() + S.idx_ptr[s];
        for (int k = 0; k < S.np[s]; ++k) I[k] = S.perm[S.first[s] + k];
        for (size_t k = 0; k < upd[s].size(); ++k) I[S.np[s] + k] = S.perm[upd[s][k]];
    }
    auto local = [&](int s, int q) -> int {   // elimination position q -> local index in front s (or -1)
        if (q >= S.first[s] && q < S.first[s] + S.np[s]) return q - S.first[s];
        const std::vector<int> &u = upd[s];
        auto it = std::lower_bound(u.begin(), u.end(), q);
        if (it == u.end() || *it != q) return -1;
        return S.np[s] + (int)(it - u.begin());
    };
    // extend-add maps
    S.rel_ptr.assign(nn + 1, 0);
    for (int s = 0; s < nn; ++s) S.rel_ptr[s + 1] = S.rel_ptr[s] + (int)upd[s].size();
    S.rel.assign(S.rel_ptr[nn], -1);
    for (int s = 0; s < nn; ++s) {
        const int p = S.parent[s];
        if (p < 0) continue;
        for (size_t k = 0; k < upd[s].size(); ++k) S.rel[S.rel_ptr[s] + k] = local(p, upd[s][k]);
    }
    // A -> front scatter map
    S.a_dest.resize(rowptr[n]);
    for (int i = 0; i < n; ++i)
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const int pi = pos[i], pj = pos[col[p]];
            const int s = node_of_pos[std::min(pi, pj)];
            const int lr = local(s, pi), lc = local(s, pj);
            S.a_dest[p] = (lr < 0 || lc < 0) ? -1 : S.front_ptr[s] + lr + (long long)lc * S.m[s];
        }
    // level schedule: level = 1 + max(children levels)
    std::vector<int> level(nn, 0);
    for (int s = 0; s < nn; ++s)
        for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) level[s] = std::max(level[s], level[S.child[ci]] + 1);
    S.nlevels = 1 + *std::max_element(level.begin(), level.end());
    S.level_ptr.assign(S.nlevels + 1, 0);
    for (int s = 0; s < nn; ++s) S.level_ptr[level[s] + 1]++;
    for (int l = 0; l < S.nlevels; ++l) S.level_ptr[l + 1] += S.level_ptr[l];
    S.level_nodes.resize(nn);
    {
        std::vector<int> next(S.level_ptr.begin(), S.level_ptr.end() - 1);
        for (int s = 0; s < nn; ++s) S.level_nodes[next[level[s]]++] = s;
    }
}

int g_mf_host_window = 1;   // pivot search window (rows): 1 = static pivoting like the device; larger = restricted partial pivoting

bool mf_factor_host(const MFSymbolic &S, const double *vals, MFHostNumeric &N) {
    N.F.assign(S.fsize, 0.0);
    N.piv.assign(S.n, 0);
    N.min_pivot = 1e300;
    for (size_t k = 0; k < S.a_dest.size(); ++k) {
        if (S.a_dest[k] < 0) return false;
        N.F[S.a_dest[k]] += vals[k];
    }
    for (int s = 0; s < S.nnodes; ++s) {      // postorder
        double *F = N.F.data() + S.front_ptr[s];
        const int m = S.m[s], np = S.np[s];
        for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
            const int c = S.child[ci], mc = S.m[c], npc = S.np[c], nu = mc - npc;
            const double *Fc = N.F.data() + S.front_ptr[c];
            const int *rel = S.rel.data() + S.rel_ptr[c];
            for (int j = 0; j < nu; ++j)
                for (int i = 0; i < nu; ++i) F[rel[i] + (size_t)rel[j] * m] += Fc[(npc + i) + (size_t)(npc + j) * mc];
        }
        for (int k = 0; k < np; ++k) {
            int p = k;
            double a = std::fabs(F[k + (size_t)k * m]);
            const int wend = (int)std::min<long long>(np, ((long long)(k / g_mf_host_window) + 1) * g_mf_host_window);
            for (int i = k + 1; i < wend; ++i)
                if (std::fabs(F[i + (size_t)k * m]) > a) {
                    a = std::fabs(F[i + (size_t)k * m]);
                    p = i;
                }
            if (a == 0.0) return false;
            N.min_pivot = std::min(N.min_pivot, a);
            N.piv[S.first[s] + k] = p;
            if (p != k)
                for (int j = 0; j < m; ++j) std::swap(F[k + (size_t)j * m], F[p + (size_t)j * m]);
            const double d = 1.0 / F[k + (size_t)k * m];
            for (int i = k + 1; i < m; ++i) F[i + (size_t)k * m] *= d;
            for (int j = k + 1; j < m; ++j) {
                const double ukj = F[k + (size_t)j * m];
                if (ukj != 0.0)
                    for (int i = k + 1; i < m; ++i) F[i + (size_t)j * m] -= F[i + (size_t)k * m] * ukj;
            }
        }
    }
    return true;
}

void mf_solve_host(const MFSymbolic &S, const MFHostNumeric &N, double *x) {
    std::vector<double> y;
    for (int s = 0; s < S.nnodes; ++s) {      // forward: leaves -> root
        const double *F = N.F.data() + S.front_ptr[s];
        const int m = S.m[s], np = S.np[s];
        const int *I = S.idx.data() + S.idx_ptr[s];
        y.resize(m);
        for (int k = 0; k < np; ++k) y[k] = x[I[k]];
        for (int k = 0; k < np; ++k) std::swap(y[k], y[N.piv[S.first[s] + k]]);
        for (int k = 0; k < np; ++k)
            for (int i = k + 1; i < np; ++i) y[i] -= F[i + (size_t)k * m] * y[k];
        for (int k = 0; k < np; ++k) x[I[k]] = y[k];
        for (int i = np; i < m; ++i) {
            double sum = 0.0;
            for (int k = 0; k < np; ++k) sum += F[i + (size_t)k * m] * y[k];
            x[I[i]] -= sum;
        }
    }
    for (int s = S.nnodes - 1; s >= 0; --s) {  // backward: root -> leaves
        const double *F = N.F.data() + S.front_ptr[s];
        const int m = S.m[s], np = S.np[s];
        const int *I = S.idx.data() + S.idx_ptr[s];
        y.resize(m);
        for (int i = 0; i < m; ++i) y[i] = x[I[i]];
        for (int k = np - 1; k >= 0; --k) {
            double sum = y[k];
            for (int j = k + 1; j < m; ++j) sum -= F[k + (size_t)j * m] * y[j];
            y[k] = sum / F[k + (size_t)k * m];
        }
        for (int k = 0; k < np; ++k) x[I[k]] = y[k];
    }
}

}  // namespace ocp
