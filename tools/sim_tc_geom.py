"""CPU model of conv_tc.cu's addressing (not of the hardware): replays tc_geom(), the TMA boxes
(with out-of-bounds zero fill), the flat 'position' indexing and the per-tap start offsets on a
NaN-initialised shared-memory image, and checks every valid output against a direct dilated
convolution.  Any valid output that depends on an unloaded byte shows up as NaN."""
import sys
import numpy as np


def round_up(a, b):
    return (a + b - 1) // b * b


def tc_geom(NKC, H, W, d):
    CP = 16 * NKC
    g = dict(H=H, W=W, d=d)
    g["side_taps"] = 1 if d < W else 0
    g["dpad"] = d if g["side_taps"] else 0
    g["Wp"] = W + g["dpad"]
    if g["Wp"] > 256:
        return None
    max_pos = min(8, 256 // CP) * 128
    Rmax = min(H, max_pos // g["Wp"])
    if Rmax < 1:
        return None
    w_bytes = 9 * NKC * 2 * CP * 16
    g["smem_w_off"] = 256 + round_up(2 * CP * 4, 128)
    g["smem_ring_off"] = round_up(g["smem_w_off"] + w_bytes, 1024)
    budget = 227 * 1024 - g["smem_ring_off"] - 4096
    best = None
    for R in range(Rmax, 0, -1):
        full, rem = divmod(H, R)
        mt = full * (-(-R * g["Wp"] // 128)) + (-(-rem * g["Wp"] // 128))
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
    g.update(R=R, tiles_per_utt=-(-H // R), n_boxes=1 if dense else 3, rows_box=(R + 2 * d + 1) if dense else (R + 1))
    g["box_stride"] = round_up(g["rows_box"] * g["Wp"] * 16, 128)
    g["slab_bytes"] = g["n_boxes"] * g["box_stride"]
    g["stage_bytes"] = 2 * g["slab_bytes"]
    g["h_start"] = [(-d if dense else (k - 1) * d) for k in range(3)]
    g["tap_off"] = [(k * d * g["Wp"] * 16 if dense else k * g["box_stride"]) for k in range(3)]
    g["n_stages"] = stages
    g["smem_total"] = max(g["smem_ring_off"] + stages * g["stage_bytes"] + 4096, 120 * 1024)
    return g


def simulate(H, W, d, NKC=1, seed=0):
    g = tc_geom(NKC, H, W, d)
    assert g is not None, (H, W, d)
    rng = np.random.default_rng(seed)
    C = 8  # one 8-channel plane is enough to exercise the addressing (K half 0)
    x = rng.standard_normal((H, W, C)).astype(np.float32)
    wt = rng.standard_normal((3, 3, C)).astype(np.float32)
    # direct reference: out[h,w] = sum_taps sum_c x[h+(dh-1)d, w+(dw-1)d, c] * wt[dh,dw,c]
    ref = np.zeros((H, W), np.float32)
    xp = np.zeros((H + 2 * d, W + 2 * d, C), np.float32)
    xp[d:d + H, d:d + W] = x
    for dh in range(3):
        for dw in range(3):
            ref += (xp[dh * d:dh * d + H, dw * d:dw * d + W] * wt[dh, dw]).sum(-1)
    out = np.full((H, W), np.nan, np.float32)
    pos_per_slab = g["slab_bytes"] // 16
    front = 4096 // 16
    for tix in range(g["tiles_per_utt"]):
        h0 = tix * g["R"]
        rows = min(g["R"], H - h0)
        n_mt = (rows * g["Wp"] + 127) // 128
        # shared memory image in units of 16-byte positions: [front pad | slab | next slab (NaN)]
        sm = np.full((front + 2 * pos_per_slab + 512, C), np.nan, np.float32)
        for bx in range(g["n_boxes"]):
            base = front + bx * g["box_stride"] // 16
            hs = h0 + g["h_start"][bx]
            for r in range(g["rows_box"]):
                for c in range(g["Wp"]):
                    hh, ww = hs + r, c - g["dpad"]
                    v = x[hh, ww] if (0 <= hh < H and 0 <= ww < W) else 0.0
                    sm[base + r * g["Wp"] + c] = v
        acc = np.zeros((n_mt * 128,), np.float32)
        for dh in range(3):
            for dw in range(3):
                if dw != 1 and not g["side_taps"]:
                    continue
                a0 = front + g["tap_off"][dh] // 16 + (dw - 1) * d
                assert g["tap_off"][dh] % 16 == 0
                for mt in range(n_mt):
                    rows_a = sm[a0 + mt * 128: a0 + mt * 128 + 128]
                    acc[mt * 128:(mt + 1) * 128] += (rows_a * wt[dh, dw]).sum(-1)
        assert (a0 + n_mt * 128) * 16 <= (front + pos_per_slab) * 16 + 4096, "over-read exceeds the tail slack"
        for pos in range(n_mt * 128):
            r, c = divmod(pos, g["Wp"])
            w = c - g["dpad"]
            if w >= 0 and r < rows:
                out[h0 + r, w] = acc[pos]
    assert not np.isnan(out).any(), f"H={H} W={W} d={d}: a valid output read unloaded shared memory"
    err = np.abs(out - ref).max()
    assert err < 1e-3, (H, W, d, err)
    return g


if __name__ == "__main__":
    cases = [(101, 40, d) for d in (1, 2, 4, 8, 16, 32, 64, 128)] + [(25, 13, 1), (50, 20, 1), (901, 40, 16),
                                                                       (301, 40, 8), (7, 5, 2), (3, 40, 1), (101, 40, 3)]
    for H, W, d in cases:
        for NKC in (1, 2, 3, 4):
            g = tc_geom(NKC, H, W, d)
            if g is None:
                print(f"H={H} W={W} d={d} NKC={NKC}: NOT TILEABLE")
                continue
            if NKC == 3:
                simulate(H, W, d, NKC)
                mt = sum((min(g['R'], H - t * g['R']) * g['Wp'] + 127) // 128 for t in range(g['tiles_per_utt']))
                print(f"H={H:4d} W={W:3d} d={d:3d}: R={g['R']:3d} tiles={g['tiles_per_utt']:3d} boxes={g['n_boxes']} "
                      f"rows_box={g['rows_box']:3d} stage={g['stage_bytes']:6d} stages={g['n_stages']} "
                      f"smem={g['smem_total']:6d} mtiles/utt={mt} eff={H * W / (mt * 128):.3f}")
    print("addressing model OK")
