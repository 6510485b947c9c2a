"""Pure-Python model of the data flow of gotoh_pair16_kernel (one alignment group,
one half of the packed registers), mirroring bgclib_b200/csrc/gotoh_pair16.cuh
statement by statement:

* lane t owns query rows [t*K, (t+1)*K) and lags lane t-1 by one step; the bottom
  Ho / F of a lane reach lane t+1 one step later (__shfl_up_sync);
* "next-T" form: the per-row state is Tn[k] = Ho[k-1][j-1] + (s - open), built with
  the profile row of the NEXT token, and E[k] already advanced to the next column;
* streaming: the group's subjects are one token stream (residues < 24, boundary 24
  behind every subject, idle 25), tokens prefetched two steps ahead; boundary /
  wind-up handling on the cold path.

Used by the CPU tests to check the kernel's recurrence, boundaries, capture and
stream logic against the oracle without a GPU.
"""
NEG = -30000
BOUNDARY = 24
IDLE = 25


def systolic_stream(q, subjects, sub, open_, ext, local, G, K):
    """Scores of query q (codes) against each subject (list of code lists)."""
    La = len(q)
    assert G * K >= La and all(len(s) > 0 for s in subjects)
    # the packed batch as the kernel sees it: idle header, then residues + >= 1 pad byte each
    mem = [IDLE] * 64
    seq_off = []
    for s in subjects:
        seq_off.append(len(mem))
        mem += list(s) + [BOUNDARY] * (16 - len(s) % 16)
    mem += [BOUNDARY] * 16
    n_g = len(subjects)
    nsteps = sum(len(s) + 1 for s in subjects) + G - 1

    def prof(c, i):              # profile word: score - open; rows past the query / token rows score 0
        return (sub[q[i]][c] if (i < La and c < 24) else 0) - open_

    def ho_init(i):              # Ho[i][-1]
        return open_ if local else 2 * open_ + i * ext

    Tn = [[0] * K for _ in range(G)]
    E = [[0] * K for _ in range(G)]
    best = [0] * G
    top_h = [open_ if local else 2 * open_] * G
    hin0 = [open_ if (local or t == 0) else 2 * open_ + (t * K - 1) * ext for t in range(G)]
    cap = [0] * G
    h_out = [0] * G
    f_out = [0] * G
    ord_ = [0] * G
    wait = [t if n_g > 0 else 0 for t in range(G)]
    ptr = [0] * G
    inc = [0] * G
    c = [IDLE] * G
    c1 = [IDLE] * G
    slots = {}
    out = [None] * n_g

    def begin_subject(t, c0):
        for k in range(K):
            above = hin0[t] if k == 0 else ho_init(t * K + k - 1)
            Tn[t][k] = above + prof(c0, t * K + k)
            E[t][k] = 0 if local else ho_init(t * K + k)
        best[t] = 0
        top_h[t] = open_ if local else 2 * open_

    if n_g > 0:
        s0 = seq_off[0]
        c[0], c1[0], ptr[0], inc[0] = mem[s0], mem[s0 + 1], s0 + 2, 1
        begin_subject(0, c[0])

    for _tau in range(nsteps):
        h_sh = [h_out[0]] + h_out[:-1]      # __shfl_up_sync(.., 1, G): lane 0 reads itself
        f_sh = [f_out[0]] + f_out[:-1]
        new_h, new_f = list(h_out), list(f_out)
        for t in range(G):
            h_in, f_in = h_sh[t], f_sh[t]
            c2 = mem[ptr[t]]
            ptr[t] += inc[t]
            if t == 0:
                h_in = top_h[t]
                f_in = 0 if local else NEG
            if c[t] < BOUNDARY:
                last = (not local) and c1[t] == BOUNDARY
                up, F = h_in, f_in
                for k in range(K):
                    i = t * K + k
                    F = max(F + ext, up)
                    h = max(Tn[t][k], E[t][k], F)
                    if local:
                        h = max(h, 0)
                        best[t] = max(best[t], h)
                    Tn[t][k] = up + prof(c1[t], i)
                    up = h + open_
                    E[t][k] = max(E[t][k] + ext, up)
                    if last and i == La - 1:
                        cap[t] = up
                new_h[t], new_f[t] = up, F
                if not local:
                    top_h[t] += ext
            elif c[t] == BOUNDARY:
                s = ord_[t]
                if local:
                    if t != G - 1:
                        slots[s] = max(slots.get(s, 0), best[t])
                    else:
                        out[s] = max(best[t], slots.pop(s, 0))
                else:
                    if t * K <= La - 1 < t * K + K:
                        out[s] = cap[t] - open_
                ord_[t] += 1
                if ord_[t] < n_g:
                    sn = seq_off[ord_[t]]
                    c1[t], c2, ptr[t] = mem[sn], mem[sn + 1], sn + 2
                    begin_subject(t, c1[t])
                else:
                    c1[t], c2, ptr[t], inc[t] = IDLE, IDLE, 0, 0
            elif wait[t] > 0:
                wait[t] -= 1
                if wait[t] == 0:
                    s0 = seq_off[0]
                    c1[t], c2, ptr[t], inc[t] = mem[s0], mem[s0 + 1], s0 + 2, 1
                    begin_subject(t, c1[t])
            c[t], c1[t] = c1[t], c2
        h_out, f_out = new_h, new_f
    return out


def systolic_score(q, s, sub, open_, ext, local, G, K):
    return systolic_stream(q, [s], sub, open_, ext, local, G, K)[0]
