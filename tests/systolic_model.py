"""Pure-Python model of the data flow of gotoh_pair16_kernel (one alignment group,
one half of the packed registers): lane t owns query rows [t*K, (t+1)*K), at step
tau it sweeps subject column j = tau - t, H/F^ of its bottom row move to lane t+1
one step later.  "Ho" form: registers keep Ho = H + open and the profile holds
s - open, as in the kernel.  Used by the CPU tests to
check the kernel's recurrence, boundaries and capture logic against the oracle
without a GPU; it mirrors bgclib_b200/csrc/gotoh_pair16.cuh statement by statement.
"""
NEG = -30000


def systolic_score(q, s, sub, open_, ext, local, G, K):
    La, Ls = len(q), len(s)
    assert G * K >= La
    Ho = [[open_ if local else 2 * open_ + (t * K + k) * ext for k in range(K)] for t in range(G)]
    E = [[0 if local else NEG for _ in range(K)] for t in range(G)]
    hin_prev = [open_ if (local or t == 0) else 2 * open_ + (t * K - 1) * ext for t in range(G)]
    h_out = [0] * G
    f_out = [0] * G
    best = 0
    cap = None
    for tau in range(Ls + G - 1):
        h_sh = [0] + h_out[:-1]          # __shfl_up_sync(.., 1, G): lane 0 keeps its own, overwritten below
        f_sh = [0] + f_out[:-1]
        new_h, new_f = list(h_out), list(f_out)
        for t in range(G):
            j = tau - t
            h_in, f_in = h_sh[t], f_sh[t]
            if t == 0:
                h_in = open_ if local else 2 * open_ + j * ext
                f_in = 0 if local else NEG
            if 0 <= j < Ls:
                c = s[j]
                diag = hin_prev[t]
                hin_prev[t] = h_in
                up, F = h_in, f_in
                for k in range(K):
                    i = t * K + k
                    prof = (sub[q[i]][c] if i < La else 0) - open_
                    E[t][k] = max(E[t][k] + ext, Ho[t][k])
                    F = max(F + ext, up)
                    h = max(diag + prof, E[t][k], F)
                    if local:
                        h = max(h, 0)
                        best = max(best, h)
                    diag = Ho[t][k]
                    up = h + open_
                    Ho[t][k] = up
                new_h[t], new_f[t] = up, F
                if not local and j == Ls - 1:
                    for k in range(K):
                        if t * K + k == La - 1:
                            cap = Ho[t][k] - open_
        h_out, f_out = new_h, new_f
    return best if local else cap
