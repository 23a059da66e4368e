// See host_lu.hpp.  Left-looking (Gilbert-Peierls) sparse LU: column k of A(:,q) is obtained by a
// sparse triangular solve with the columns of L found so far; the non-zero set of that solve is the
// reach of the column's pattern in the graph of L (depth-first search, topological order).
#include "host_lu.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace ocp {

namespace {

struct NDState {
    int n;
    const int *rowptr, *col;
    const double *xy;
    const uint8_t *kind;
    std::vector<int> side;   // scratch: 0 outside current set, 1 left, 2 right
    std::vector<int> *out;
    int leaf;
};

void nd_emit(NDState &s, std::vector<int> &set) {
    std::stable_sort(set.begin(), set.end(), [&](int a, int b) {
        if (s.kind[a] != s.kind[b]) return s.kind[a] < s.kind[b];
        return a < b;
    });
    s.out->insert(s.out->end(), set.begin(), set.end());
}

void nd_rec(NDState &s, std::vector<int> &set) {
    if ((int)set.size() <= s.leaf) {
        nd_emit(s, set);
        return;
    }
    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
    for (int i : set)
        for (int d = 0; d < 2; ++d) {
            lo[d] = std::min(lo[d], s.xy[2 * i + d]);
            hi[d] = std::max(hi[d], s.xy[2 * i + d]);
        }
    int ax = (hi[0] - lo[0] >= hi[1] - lo[1]) ? 0 : 1;
    std::vector<double> c(set.size());
    for (size_t k = 0; k < set.size(); ++k) c[k] = s.xy[2 * set[k] + ax];
    std::nth_element(c.begin(), c.begin() + c.size() / 2, c.end());
    double med = c[c.size() / 2];
    std::vector<int> left, right, sep;
    for (int i : set) {
        if (s.xy[2 * i + ax] < med) {
            s.side[i] = 1;
            left.push_back(i);
        } else {
            s.side[i] = 2;
            right.push_back(i);
        }
    }
    if (left.empty() || right.empty()) {
        for (int i : set) s.side[i] = 0;
        nd_emit(s, set);
        return;
    }
    std::vector<int> rest;
    for (int i : right) {
        bool touches = false;
        for (int p = s.rowptr[i]; p < s.rowptr[i + 1] && !touches; ++p) touches = (s.side[s.col[p]] == 1);
        (touches ? sep : rest).push_back(i);
    }
    for (int i : set) s.side[i] = 0;
    set.clear();
    set.shrink_to_fit();
    nd_rec(s, left);
    nd_rec(s, rest);
    // order the separator along the cut so that its dense block is banded
    std::stable_sort(sep.begin(), sep.end(), [&](int a, int b) {
        if (s.kind[a] != s.kind[b]) return s.kind[a] < s.kind[b];
        return s.xy[2 * a + (1 - ax)] < s.xy[2 * b + (1 - ax)];
    });
    s.out->insert(s.out->end(), sep.begin(), sep.end());
}

}  // namespace

void nested_dissection_order(int n, const int *rowptr, const int *col, const double *xy, const uint8_t *kind,
                             std::vector<int> &q) {
    NDState s{n, rowptr, col, xy, kind, std::vector<int>(n, 0), &q, 48};
    q.clear();
    q.reserve(n);
    std::vector<int> all(n);
    std::iota(all.begin(), all.end(), 0);
    nd_rec(s, all);
}

bool sparse_lu(int n, const int *rowptr, const int *col, const double *val, const std::vector<int> &q,
               double thresh, HostLU &out) {
    const int nnz = rowptr[n];
    // CSC copy of A
    std::vector<int> Ap(n + 1, 0), Ai(nnz);
    std::vector<double> Ax(nnz);
    for (int p = 0; p < nnz; ++p) Ap[col[p] + 1]++;
    for (int j = 0; j < n; ++j) Ap[j + 1] += Ap[j];
    {
        std::vector<int> next(Ap.begin(), Ap.end() - 1);
        for (int i = 0; i < n; ++i)
            for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
                int d = next[col[p]]++;
                Ai[d] = i;
                Ax[d] = val[p];
            }
    }
    // factors, column-wise, row indices in ORIGINAL numbering until the end
    std::vector<int> Lp(n + 1, 0), Up(n + 1, 0), Li, Ui;
    std::vector<double> Lx, Ux;
    Li.reserve((size_t)nnz * 4);
    Lx.reserve((size_t)nnz * 4);
    Ui.reserve((size_t)nnz * 4);
    Ux.reserve((size_t)nnz * 4);
    std::vector<int> pinv(n, -1), xi(n), stack(n), pstack(n), mark(n, -1);
    std::vector<double> x(n, 0.0);
    double minpiv = 1e300, maxpiv = 0.0;

    for (int k = 0; k < n; ++k) {
        const int c = q[k];
        Lp[k] = (int)Li.size();
        Up[k] = (int)Ui.size();
        // ---- reach of A(:,c) in the graph of L: non-recursive DFS, topological order in xi[top..n)
        int top = n;
        for (int p = Ap[c]; p < Ap[c + 1]; ++p) {
            int r = Ai[p];
            if (mark[r] == k) continue;
            int head = 0;
            stack[0] = r;
            while (head >= 0) {
                int j = stack[head];
                int jn = pinv[j];
                if (mark[j] != k) {
                    mark[j] = k;
                    pstack[head] = (jn < 0) ? 0 : Lp[jn] + 1;   // skip the unit diagonal
                }
                bool done = true;
                int pend = (jn < 0) ? 0 : Lp[jn + 1];
                for (int p2 = pstack[head]; p2 < pend; ++p2) {
                    int i = Li[p2];
                    if (mark[i] == k) continue;
                    pstack[head] = p2 + 1;
                    stack[++head] = i;
                    done = false;
                    break;
                }
                if (done) {
                    --head;
                    xi[--top] = j;
                }
            }
        }
        for (int p = top; p < n; ++p) x[xi[p]] = 0.0;
        for (int p = Ap[c]; p < Ap[c + 1]; ++p) x[Ai[p]] = Ax[p];
        // ---- sparse triangular solve
        for (int px = top; px < n; ++px) {
            int j = xi[px], jn = pinv[j];
            if (jn < 0) continue;
            double xj = x[j];
            for (int p = Lp[jn] + 1; p < Lp[jn + 1]; ++p) x[Li[p]] -= Lx[p] * xj;
        }
        // ---- pivot
        int ipiv = -1;
        double a = -1.0;
        for (int p = top; p < n; ++p) {
            int i = xi[p];
            if (pinv[i] < 0) {
                double t = std::fabs(x[i]);
                if (t > a) {
                    a = t;
                    ipiv = i;
                }
            } else {
                Ui.push_back(pinv[i]);
                Ux.push_back(x[i]);
            }
        }
        if (ipiv < 0 || a <= 0.0) return false;
        if (pinv[c] < 0 && mark[c] == k && std::fabs(x[c]) >= thresh * a) ipiv = c;
        double pivot = x[ipiv];
        minpiv = std::min(minpiv, std::fabs(pivot));
        maxpiv = std::max(maxpiv, std::fabs(pivot));
        Ui.push_back(k);
        Ux.push_back(pivot);
        pinv[ipiv] = k;
        Li.push_back(ipiv);
        Lx.push_back(1.0);
        for (int p = top; p < n; ++p) {
            int i = xi[p];
            if (pinv[i] < 0) {
                Li.push_back(i);
                Lx.push_back(x[i] / pivot);
            }
            x[i] = 0.0;
        }
    }
    Lp[n] = (int)Li.size();
    Up[n] = (int)Ui.size();
    for (size_t p = 0; p < Li.size(); ++p) Li[p] = pinv[Li[p]];

    // ---- CSC -> CSR (columns come out ascending)
    auto to_csr = [n](const std::vector<int> &Cp, const std::vector<int> &Ci, const std::vector<double> &Cx,
                      std::vector<int> &Rp, std::vector<int> &Rj, std::vector<double> &Rx) {
        Rp.assign(n + 1, 0);
        Rj.resize(Ci.size());
        Rx.resize(Ci.size());
        for (int i : Ci) Rp[i + 1]++;
        for (int i = 0; i < n; ++i) Rp[i + 1] += Rp[i];
        std::vector<int> next(Rp.begin(), Rp.end() - 1);
        for (int j = 0; j < n; ++j)
            for (int p = Cp[j]; p < Cp[j + 1]; ++p) {
                int d = next[Ci[p]]++;
                Rj[d] = j;
                Rx[d] = Cx[p];
            }
    };
    out.n = n;
    to_csr(Lp, Li, Lx, out.Lp, out.Li, out.Lx);
    to_csr(Up, Ui, Ux, out.Up, out.Ui, out.Ux);
    out.P.assign(n, 0);
    for (int i = 0; i < n; ++i) out.P[pinv[i]] = i;
    out.Q = q;
    out.min_pivot = minpiv;
    out.max_pivot = maxpiv;
    return true;
}

void host_lu_solve(const HostLU &lu, double *b) {
    const int n = lu.n;
    std::vector<double> y(n);
    for (int i = 0; i < n; ++i) y[i] = b[lu.P[i]];
    for (int i = 0; i < n; ++i) {          // L y = Pb   (unit diagonal stored last in each row)
        double s = y[i];
        for (int p = lu.Lp[i]; p < lu.Lp[i + 1]; ++p)
            if (lu.Li[p] < i) s -= lu.Lx[p] * y[lu.Li[p]];
        y[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {     // U z = y     (diagonal first in each row)
        double s = y[i], d = 0.0;
        for (int p = lu.Up[i]; p < lu.Up[i + 1]; ++p) {
            if (lu.Ui[p] == i) d = lu.Ux[p];
            else s -= lu.Ux[p] * y[lu.Ui[p]];
        }
        y[i] = s / d;
    }
    for (int j = 0; j < n; ++j) b[lu.Q[j]] = y[j];
}

}  // namespace ocp
