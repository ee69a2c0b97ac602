#!/usr/bin/env python3
"""Per-tree time stamps of the MCTS wave kernels (spl_mcts_debug_profile): where inside a wave the time goes.

    python profiles/tools/wave_stamps.py [--trees 4096] [--sims 1600] [--moves 3]

Brings configs[1] to its steady state (asynchronous self-play), then runs plain (non-graph) overlapped waves with the
profile buffer attached and prints, per kernel: launch-to-last-warp duration (globaltimer), the distribution of the
per-warp durations (SM clock) and of the phases inside them. Diagnostics only.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import azg_b200 as azg   # noqa: E402
from azg_b200 import _native as nat   # noqa: E402


def pct(x):
    x = np.asarray(x, dtype=np.float64)
    return {"mean": float(x.mean()), "p50": float(np.percentile(x, 50)), "p90": float(np.percentile(x, 90)), "p99": float(np.percentile(x, 99)),
            "max": float(x.max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trees", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=1600)
    ap.add_argument("--moves", type=float, default=3.0)
    ap.add_argument("--waves", type=int, default=40)
    ap.add_argument("--max-levels", type=int, default=16)
    a = ap.parse_args()
    n, T = 2, a.trees
    net = azg.FusedSplendorNNet(n, seed=1, device=0)
    cap = 8 * a.sims
    eng = azg.SelfPlayEngine(n, T, net, a.sims, device=0, seed=1, node_cap=cap, edge_cap=cap * 36, gc_reachable=True, graph_waves=64,
                             max_levels=a.max_levels, clean_every=max(1, int(4 * a.sims / 64)), clean_percent=45)
    eng.env.rollout(24, rotate=True)
    eng.start_async()
    for _ in range(int(a.moves * a.sims / 64)):
        eng.tick(64)
    ar = eng.arena
    prof = torch.zeros((T, 16), dtype=torch.int64, device=ar.device)
    nat.check(ar._lib.spl_mcts_debug_profile(ar._m, C.c_void_p(prof.data_ptr())))
    mhz = 1965.0
    acc = {}
    for w in range(a.waves):
        prof.zero_()
        ar.wave_nnet(net)
        torch.cuda.synchronize()
        p = prof.cpu().numpy()
        if w < 5:
            continue
        ed_us = (p[:, 3] - p[:, 1]) / mhz
        rw = p[:, 6] > 1e15      # (the separate rules kernel stamps a globaltimer there; the fused path keeps small counters in these slots)
        at = p[:, 13] != 0
        cur = {
            "ed_kernel_span_us": (p[:, 5].max() - p[:, 0].min()) / 1e3,
            "ed_warp_start_spread_us": (p[:, 0].max() - p[:, 0].min()) / 1e3,
            "ed_warp_us": pct(ed_us), "ed_expand_us": pct((p[:, 2] - p[:, 1]) / mhz), "ed_descend_us": pct((p[:, 3] - p[:, 2]) / mhz),
            "ed_levels": pct(p[:, 4] % 1000),
            "ed_us_per_level_p50": float(np.median(((p[:, 3] - p[:, 2]) / mhz)[(p[:, 4] % 1000) >= 4] / (p[:, 4] % 1000)[(p[:, 4] % 1000) >= 4])),
            "attach_kernel_span_us": (p[at, 15].max() - p[at, 12].min()) / 1e3,
            "attach_warp_us": pct((p[at, 14] - p[at, 13]) / mhz),
        }
        if rw.any():      # separate rules kernel
            cur.update({
                "rules_kernel_span_us": (p[rw, 11].max() - p[rw, 6].min()) / 1e3,
                "rules_warp_us": pct((p[rw, 10] - p[rw, 7]) / mhz), "rules_load_us": pct((p[rw, 8] - p[rw, 7]) / mhz),
                "rules_core_us": pct((p[rw, 9] - p[rw, 8]) / mhz), "rules_store_us": pct((p[rw, 10] - p[rw, 9]) / mhz),
                "rules_start_after_ed_end_us": (p[rw, 6].min() - p[:, 5].max()) / 1e3,
                "attach_start_after_rules_end_us": (p[at, 12].min() - p[rw, 11].max()) / 1e3,
            })
        else:             # rules step inside the descent kernel
            lv = p[:, 7] > 0
            cur["ed_pick_us_per_level"] = pct((p[lv, 6] / p[lv, 7]) / mhz)
            cur["ed_pick_share_of_descent"] = float((p[lv, 6] / mhz).sum() / ((p[lv, 3] - p[lv, 2]) / mhz).sum())
            fr = (p[:, 4] // 1000) == 2
            fr = fr & (p[:, 8] != 0)
            cur.update({"rules_load_us": pct((p[fr, 8] - p[fr, 3]) / mhz), "rules_move_and_rotate_us": pct((p[fr, 10] - p[fr, 8]) / mhz),
                        "rules_card_flags_us": pct((p[fr, 11] - p[fr, 10]) / mhz), "rules_ended_mask_store_us": pct((p[fr, 9] - p[fr, 11]) / mhz)})
            cur.update({"ed_rules_step_us": pct((p[fr, 9] - p[fr, 3]) / mhz), "ed_warp_with_rules_us": pct((p[:, 9] - p[:, 1]) / mhz),
                        "attach_start_after_ed_end_us": (p[at, 12].min() - p[:, 5].max()) / 1e3})
        for k, v in cur.items():
            acc.setdefault(k, []).append(v)
    out = {}
    for k, v in acc.items():
        if isinstance(v[0], dict):
            out[k] = {kk: round(float(np.mean([x[kk] for x in v])), 2) for kk in v[0]}
        else:
            out[k] = round(float(np.mean(v)), 2)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
