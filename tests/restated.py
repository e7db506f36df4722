"""Pure-Python model of the RESTATED stem recurrence that the CUDA kernel implements:
non-leaf nodes only, K tables replaced by root->node path counts, MATCH folded into the y-sweep
through the row-local sums Q.  Small cases only; used by tests to show that the restatement equals
the literal reference recurrence (oracle) before any GPU is involved."""
import math

import numpy as np


def compile_record(f, g):
    """f: MData.export() dict.  Mirrors stem_kernel_b200/csrc/compile_set.cpp."""
    n = len(f["first"])
    eoff, eto, gaps, ew = f["edge_off"], f["edge_to"], f["edge_gaps"], f["edge_w"]
    gp = [1.0]
    for _ in range(int(max(gaps, default=0))):
        gp.append(gp[-1] * g)
    leaf = [eoff[u] == eoff[u + 1] for u in range(n)]
    px, ql, el, av, pl = [0.0] * n, [0.0] * n, [0.0] * n, [0.0] * n, [0.0] * n
    for u in range(n):
        if leaf[u]:
            px[u] = pl[u] = 1.0
            continue
        av[u] = g * g * float(f["weight"][u])
        q = e_leaf = p_l = 0.0
        for e in range(eoff[u], eoff[u + 1]):
            c = int(eto[e])
            ce = gp[int(gaps[e])] * float(ew[e])
            q += ce * px[c]
            p_l += pl[c]
            if leaf[c]:
                e_leaf += ce
        ql[u], el[u], px[u], pl[u] = q, e_leaf, av[u] * q, p_l
    paths = [0.0] * n
    plr, lr = 0.0, 0
    for u in f["root"]:
        paths[int(u)] += 1.0
        plr += pl[int(u)]
        lr += leaf[int(u)]
    for u in range(n - 1, -1, -1):
        for e in range(eoff[u], eoff[u + 1]):
            paths[int(eto[e])] += paths[u]
    nl = [u for u in range(n) if not leaf[u]]
    new = {u: k for k, u in enumerate(nl)}
    ch = [[(new[int(eto[e])], gp[int(gaps[e])] * float(ew[e])) for e in range(eoff[u], eoff[u + 1])
           if not leaf[int(eto[e])]] for u in nl]
    bpf = [[(int(f["bpf_a"][b]) * 4 + int(f["bpf_b"][b]), float(f["bpf_f"][b]))
            for b in range(f["bpf_off"][u], f["bpf_off"][u + 1])] for u in nl]
    gapt = [float(f["profile"][int(f["first"][u])][4]) / f["n_seqs"] for u in nl]
    return dict(N=len(nl), a=[av[u] for u in nl], el=[el[u] for u in nl], ql=[ql[u] for u in nl],
                paths=[paths[u] for u in nl], length=[int(f["last"][u]) - int(f["first"][u]) for u in nl], ch=ch,
                bpf=bpf, gapt=gapt, plr=plr, lr=lr)


def stem_pair(x, y, pair_tab, len_band):
    """x, y: compile_record() outputs; pair_tab: 256 doubles (index ab*16+cd)."""
    Nx, Ny = x["N"], y["N"]
    G0 = np.zeros((Nx, Ny))
    total = 0.0
    for i in range(Nx):                      # any children-first order works
        Q = np.zeros(Ny)
        for c, e in x["ch"][i]:
            Q += e * G0[c]
        G1 = np.zeros(Ny)
        row = 0.0
        for j in range(Ny):
            R = y["el"][j] * x["ql"][i]
            S = 0.0
            for c, e in y["ch"][j]:
                R += e * Q[c]
                S += e * G1[c]
            m = 0.0
            if len_band == 0 or abs(x["length"][i] - y["length"][j]) <= len_band:
                vs = 0.0
                for ab, fx in x["bpf"][i]:
                    for cd, fy in y["bpf"][j]:
                        vs += pair_tab[ab * 16 + cd] * fx * fy
                vs += y["a"][j] * x["gapt"][i] + x["a"][i] * y["gapt"][j]
                m = vs * R
            G1[j] = m + y["a"][j] * S
            G0[i, j] = G1[j] + x["a"][i] * Q[j]
            row += y["paths"][j] * m
        total += x["paths"][i] * row
    return total + x["plr"] * y["lr"]


def string_pair_sum(sx, sy, wx, wy, subst, gap):
    """K-free form of the lite string kernel: K0[Lx][Ly] = 1 + sum of all MATCH terms v(i,j)."""
    Lx, Ly = len(sx), len(sy)
    G0 = [[gap ** 0] * (Ly + 1) for _ in range(Lx + 1)]
    for j in range(1, Ly + 1):
        G0[0][j] = G0[0][j - 1] * gap
    total = 1.0
    for i in range(1, Lx + 1):
        G0[i][0] = G0[i - 1][0] * gap
        g1 = 0.0
        for j in range(1, Ly + 1):
            v = G0[i - 1][j - 1] * wx[i - 1] * wy[j - 1] * subst[sx[i - 1] * 4 + sy[j - 1]]
            total += v
            g1 = v + g1 * gap
            G0[i][j] = g1 + G0[i - 1][j] * gap
    return total
