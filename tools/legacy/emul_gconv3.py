"""CPU emulation of gconv3 (csrc/gconv3.cu): (1) host geometry + kernel index arithmetic -- every output pixel is
written exactly once, every A-operand read stays inside the TMA box, shift-add partners stay inside the M-tile;
(2) numpy emulation of the MMAs (A start, B rows, N per tap group, first-MMA overwrite) and of the shift-add epilogue
against a direct convolution.  Validates the algebra, not the hardware mechanics.  python tools/emul_gconv3.py"""


import itertools
def check(H, W, k, Cout, KC=64, issuers=2, smem_lim=210*1024):
    TPM = 128 // Cout; MS = 128 - (TPM - 1)
    Wp = W + k - 1; npos = H * Wp
    for mt_try in range(issuers, 0, -1):
        PT = mt_try * MS
        n = (npos + PT - 1) // PT
        box_rows = 0
        for j in range(n):
            p0 = j * PT; c0 = p0 % Wp
            mt_n = min(mt_try, (npos - p0 + MS - 1) // MS)
            rows = (c0 + (mt_n - 1) * MS + 127 + (k - 1) * (Wp + 1)) // Wp + 1
            box_rows = max(box_rows, rows)
        box_bytes = box_rows * Wp * KC * 2
        if 2 * ((box_bytes + 1023) // 1024 * 1024) + 4 * 128 * KC * 2 + 1024 <= smem_lim:
            break
    else:
        return f"H={H} W={W} k={k} Cout={Cout}: does not fit"
    written = {}
    G = (k + TPM - 1) // TPM
    for j in range(n):
        p0 = j * PT; h0 = p0 // Wp; c0 = p0 - h0 * Wp
        mt_n = min(mt_try, (npos - p0 + MS - 1) // MS)
        assert mt_n >= 1
        for mt in range(mt_n):
            for L in range(128):
                # A reads of row L for all taps: box index
                for tr in range(k):
                    for g in range(G):
                        idx = c0 + mt * MS + L + tr * Wp + g * TPM
                        assert idx < box_rows * Wp, ("A read outside box", H, W, k, Cout, j, mt, L, tr, g, idx, box_rows * Wp)
                pa = p0 + mt * MS + L
                hl, w = divmod(pa, Wp)
                valid = L < MS and hl < H and w < W
                if valid:
                    nblk = min(k, TPM)
                    assert L + nblk - 1 <= 127
                    key = (hl, w)
                    written[key] = written.get(key, 0) + 1
                    # semantic check of the tap mapping: block jj of row L+jj holds tap s0+jj at position pa+jj evaluated with
                    # input offset s0 -> input column (w + jj) + s0 - pad ... equals w + (s0 + jj) - pad: OK by construction
    missing = [(h, w) for h in range(H) for w in range(W) if written.get((h, w), 0) != 1]
    assert not missing, ("coverage", H, W, k, Cout, missing[:5])
    return f"H={H} W={W} k={k} Cout={Cout}: mt/tile {mt_try}, tiles/sample {n}, box rows {box_rows} ({box_bytes/1024:.1f} KB), M-tiles/sample {sum(min(mt_try, (npos - j*PT + MS - 1)//MS) for j in range(n))} (ideal {H*W/128:.2f})"
for (H, W) in [(32, 32), (16, 16), (64, 64), (8, 8), (20, 12), (9, 40), (24, 20), (1, 1), (255, 3)]:
    for k in (1, 3, 5, 7):
        for Cout in (32, 64):
            print(check(H, W, k, Cout))


import numpy as np  # noqa: E402
def run(H, W, k, Cout, Cin=8, mt_per_tile=2):
    rng = np.random.default_rng(H * 100 + W * 10 + k + Cout)
    TPM = 128 // Cout; MS = 128 - (TPM - 1); pad = (k - 1) // 2
    Wp = W + k - 1; npos = H * Wp; PT = mt_per_tile * MS
    X = rng.standard_normal((H, W, Cin)); Wt = rng.standard_normal((k * k + 8, Cout, Cin))   # extra rows = next expert's taps (garbage)
    n = (npos + PT - 1) // PT
    Y = np.full((H, W, Cout), np.nan)
    G = (k + TPM - 1) // TPM
    nblk = min(k, TPM)
    for j in range(n):
        p0 = j * PT; h0 = p0 // Wp; c0 = p0 - h0 * Wp
        mt_n = min(mt_per_tile, (npos - p0 + MS - 1) // MS)
        rows = (c0 + (mt_n - 1) * MS + 127 + (k - 1) * (Wp + 1)) // Wp + 1
        box = np.zeros((rows, Wp, Cin))                       # TMA box with OOB zero fill, origin (h0 - pad, -pad)
        for rr in range(rows):
            for cc in range(Wp):
                hh, ww = h0 - pad + rr, cc - pad
                if 0 <= hh < H and 0 <= ww < W: box[rr, cc] = X[hh, ww]
        flat = box.reshape(rows * Wp, Cin)
        for mt in range(mt_n):
            D = np.zeros((128, 128))
            first = True
            for tr in range(k):
                for g in range(G):
                    ntaps = min(TPM, k - g * TPM)
                    A = flat[c0 + mt * MS + tr * Wp + g * TPM: c0 + mt * MS + tr * Wp + g * TPM + 128]      # [128][Cin]
                    B = Wt.reshape(-1, Cin)[(tr * k + g * TPM) * Cout: (tr * k + g * TPM) * Cout + ntaps * Cout]   # [N][Cin]
                    P = A @ B.T
                    if first: D[:, :ntaps * Cout] = P; first = False
                    else: D[:, :ntaps * Cout] += P
            for L in range(MS):
                pa = p0 + mt * MS + L
                hl, w = divmod(pa, Wp)
                if hl < H and w < W:
                    Y[hl, w] = sum(D[L + jj, jj * Cout:(jj + 1) * Cout] for jj in range(nblk))
    ref = np.zeros((H, W, Cout))
    for h in range(H):
        for w in range(W):
            for tr in range(k):
                for ts in range(k):
                    hh, ww = h + tr - pad, w + ts - pad
                    if 0 <= hh < H and 0 <= ww < W: ref[h, w] += Wt[tr * k + ts] @ X[hh, ww]
    err = np.abs(Y - ref).max()
    print(f"H={H} W={W} k={k} Cout={Cout}: max abs err {err:.2e}", "OK" if err < 1e-9 else "MISMATCH")
for args in [(32, 32, 5, 64), (32, 32, 3, 64), (16, 16, 5, 64), (16, 16, 5, 32), (16, 16, 3, 32), (12, 10, 7, 32), (9, 20, 1, 64), (8, 8, 7, 64), (20, 12, 1, 32)]:
    run(*args)
