"""CPU model of conv_tc.cu's addressing (not of the hardware): replays tc_geom(), the TMA boxes
(with out-of-bounds zero fill and, in phase mode, the strided row dimension), the flat
'position' indexing and the per-tap start offsets on a NaN-initialised shared-memory image, and
checks every valid output against a direct dilated convolution.  Any valid output that depends
on an unloaded byte shows up as NaN."""
import numpy as np


def round_up(a, b):
    return (a + b - 1) // b * b


def cdiv(a, b):
    return -(-a // b)


def use_phase(H, d):
    return 8 <= d <= 32 and 2 * d <= H


def tc_geom(NKC, H, Hpad, W, d):
    CP = 16 * NKC
    g = dict(H=H, W=W, d=d, Hpad=Hpad)
    g["side_taps"] = 1 if d < W else 0
    g["dpad"] = d if g["side_taps"] else 0
    g["Wp"] = W + g["dpad"]
    if g["Wp"] > 256:
        return None
    max_pos = min(8, 256 // CP) * 128
    w_bytes = 9 * NKC * 2 * CP * 16
    g["smem_w_off"] = 256 + round_up(2 * CP * 4, 128)
    g["smem_ring_off"] = round_up(g["smem_w_off"] + w_bytes, 1024)
    budget = 227 * 1024 - g["smem_ring_off"] - 4096
    g["phase"] = 1 if (use_phase(H, d) and Hpad % d == 0) else 0
    g["chunks_per_phase"] = 1
    if g["phase"]:
        rows_max = cdiv(H, d)
        Rmax = min(rows_max, max_pos // g["Wp"])
        if Rmax < 1:
            return None
        best = None
        for R in range(Rmax, 0, -1):
            mt = 0
            for ph in range(d):
                n = cdiv(H - ph, d)
                mt += (n // R) * cdiv(R * g["Wp"], 128) + cdiv((n % R) * g["Wp"], 128)
            if best is None or mt < best[1]:
                best = (R, mt)
        R = best[0]
        g.update(R=R, chunks_per_phase=cdiv(rows_max, R), n_boxes=1, rows_box=R + 3)
        g["tiles_per_utt"] = d * g["chunks_per_phase"]
        g["box_stride"] = round_up(g["rows_box"] * g["Wp"] * 16, 128)
        g["slab_bytes"] = g["box_stride"]
        g["stage_bytes"] = 2 * g["slab_bytes"]
        g["h_start"] = [-1, -1, -1]
        g["tap_off"] = [k * g["Wp"] * 16 for k in range(3)]
        g["n_stages"] = min(8, budget // g["stage_bytes"])
        if g["n_stages"] < 2:
            return None
    else:
        Rmax = min(H, max_pos // g["Wp"])
        if Rmax < 1:
            return None
        best = None
        for R in range(Rmax, 0, -1):
            full, rem = divmod(H, R)
            mt = full * cdiv(R * g["Wp"], 128) + cdiv(rem * g["Wp"], 128)
            if best is not None and mt >= best[1]:
                continue
            dense = d <= R
            rows_box = R + 2 * d + 1 if dense else R + 1
            if rows_box > 256:
                continue
            slab = (1 if dense else 3) * round_up(rows_box * g["Wp"] * 16, 128)
            if slab >= (1 << 18):
                continue
            stages = min(8, budget // (2 * slab))
            if stages < 2:
                continue
            best = (R, mt, stages)
        if best is None:
            return None
        R, _, stages = best
        dense = d <= R
        g.update(R=R, tiles_per_utt=cdiv(H, R), n_boxes=1 if dense else 3, rows_box=(R + 2 * d + 1) if dense else (R + 1))
        g["box_stride"] = round_up(g["rows_box"] * g["Wp"] * 16, 128)
        g["slab_bytes"] = g["n_boxes"] * g["box_stride"]
        g["stage_bytes"] = 2 * g["slab_bytes"]
        g["h_start"] = [(-d if dense else (k - 1) * d) for k in range(3)]
        g["tap_off"] = [(k * d * g["Wp"] * 16 if dense else k * g["box_stride"]) for k in range(3)]
        g["n_stages"] = stages
    g["smem_total"] = max(g["smem_ring_off"] + g["n_stages"] * g["stage_bytes"] + 4096, 120 * 1024)
    return g


def tile_decode(g, tix):
    if g["phase"]:
        ph = tix // g["chunks_per_phase"]
        r0 = (tix - ph * g["chunks_per_phase"]) * g["R"]
        rows = min(g["R"], cdiv(g["H"] - ph, g["d"]) - r0)
    else:
        ph, r0 = 0, tix * g["R"]
        rows = min(g["R"], g["H"] - r0)
    return ph, r0, rows


def simulate(H, W, d, NKC=3, seed=0, Hpad=None):
    Hpad = Hpad or (round_up(H, d) if use_phase(H, d) else H)
    g = tc_geom(NKC, H, Hpad, W, d)
    assert g is not None, (H, W, d)
    rng = np.random.default_rng(seed)
    C = 8  # one 8-channel plane is enough to exercise the addressing
    x = rng.standard_normal((H, W, C)).astype(np.float32)
    mem = np.zeros((Hpad, W, C), np.float32)   # plane in memory, pad rows zero
    mem[:H] = x
    wt = rng.standard_normal((3, 3, C)).astype(np.float32)
    ref = np.zeros((H, W), np.float32)
    xp = np.zeros((H + 2 * d, W + 2 * d, C), np.float32)
    xp[d:d + H, d:d + W] = x
    for dh in range(3):
        for dw in range(3):
            ref += (xp[dh * d:dh * d + H, dw * d:dw * d + W] * wt[dh, dw]).sum(-1)
    out = np.full((H, W), np.nan, np.float32)
    pos_per_slab = g["slab_bytes"] // 16
    front = 4096 // 16
    hstep = d if g["phase"] else 1
    n_rows_dim = (Hpad // d) if g["phase"] else H     # tensor-map extent of the row dimension
    total_mt = 0
    for tix in range(g["tiles_per_utt"]):
        ph, r0, rows = tile_decode(g, tix)
        if rows <= 0:
            continue
        n_mt = cdiv(rows * g["Wp"], 128)
        total_mt += n_mt
        sm = np.full((front + 2 * pos_per_slab + 512, C), np.nan, np.float32)
        for bx in range(g["n_boxes"]):
            base = front + bx * g["box_stride"] // 16
            rs = r0 + g["h_start"][bx]
            for r in range(g["rows_box"]):
                for c in range(g["Wp"]):
                    rr, ww = rs + r, c - g["dpad"]
                    if 0 <= rr < n_rows_dim and 0 <= ww < W:
                        v = mem[rr * hstep + ph, ww]       # in-bounds for the tensor map: real memory (pad rows are 0)
                    else:
                        v = 0.0                            # TMA out-of-bounds fill
                    sm[base + r * g["Wp"] + c] = v
        acc = np.zeros((n_mt * 128,), np.float32)
        for dh in range(3):
            for dw in range(3):
                if dw != 1 and not g["side_taps"]:
                    continue
                assert g["tap_off"][dh] % 16 == 0
                a0 = front + g["tap_off"][dh] // 16 + (dw - 1) * d
                for mt in range(n_mt):
                    acc[mt * 128:(mt + 1) * 128] += (sm[a0 + mt * 128: a0 + mt * 128 + 128] * wt[dh, dw]).sum(-1)
        assert (a0 + n_mt * 128) * 16 <= (front + pos_per_slab) * 16 + 4096, "over-read exceeds the tail slack"
        for pos in range(n_mt * 128):
            r, c = divmod(pos, g["Wp"])
            w = c - g["dpad"]
            if w >= 0 and r < rows:
                h = (r0 + r) * hstep + ph
                assert h < H
                assert np.isnan(out[h, w]), "pixel written twice"
                out[h, w] = acc[pos]
    assert not np.isnan(out).any(), f"H={H} W={W} d={d}: a pixel is missing or read unloaded shared memory"
    err = np.abs(out - ref).max()
    assert err < 1e-3, (H, W, d, err)
    g["total_mt"] = total_mt
    return g


if __name__ == "__main__":
    cases = [(101, 40, d) for d in (1, 2, 4, 8, 16, 32, 64, 128)] + [(25, 13, 1), (50, 20, 1), (901, 40, 16), (901, 40, 32),
                                                                       (301, 40, 8), (7, 5, 2), (3, 40, 1), (101, 40, 3),
                                                                       (301, 40, 16), (37, 40, 8), (16, 40, 8)]
    for H, W, d in cases:
        for NKC in (1, 2, 3, 4):
            Hpad = round_up(H, 16) if use_phase(H, d) else H
            if tc_geom(NKC, H, Hpad, W, d) is None:
                print(f"H={H} W={W} d={d} NKC={NKC}: NOT TILEABLE")
                continue
            if NKC in (2, 3):
                g = simulate(H, W, d, NKC, Hpad=Hpad)
                if NKC == 3:
                    print(f"H={H:4d} W={W:3d} d={d:3d}: phase={g['phase']} R={g['R']:3d} tiles={g['tiles_per_utt']:3d} "
                          f"boxes={g['n_boxes']} rows_box={g['rows_box']:3d} stage={g['stage_bytes']:6d} stages={g['n_stages']} "
                          f"smem={g['smem_total']:6d} mtiles/utt={g['total_mt']} eff={H * W / (g['total_mt'] * 128):.3f} "
                          f"halo x{g['tiles_per_utt'] * g['n_boxes'] * g['rows_box'] / H:.2f}")
    print("addressing model OK")
