"""A/B on one B200: the reference's OWN raymarching kernels (oracle/_ref/_raymarching.so = the unmodified
submodules/raymarching/src/raymarching.cu rebuilt for sm_100a by oracle/build_ref.sh) next to this repo's kernels for the
same operator, same inputs -- BASELINE.md section 3 "B-ref-GPU", SURVEY section 2a "the bar for each is the reference
kernel itself recompiled for sm_100a on the same box".

Each operator is timed as its reference WRAPPER runs it (submodules/raymarching/raymarching.py): the wrapper's zero
fills are part of the reference's cost of the op (e.g. march_rays_train zero-fills xyzs/dirs/deltas every call,
raymarching.py:205-207) and are issued here with the same torch calls; ours needs none.  CUDA events on the launching
stream, 5 warm-up + 20 timed calls back to back.  Sizes: cfg2 (4096 rays), 2^18 rays, cfg3 (640 000 rays, first loop
iteration).  Runs as a subprocess of bench.py (the reference extension launches on the legacy default stream and links
libtorch: kept out of the bench process) and prints one JSON object.

    python scripts/ref_gpu_kernels.py [--sizes 4096,262144,640000]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

WARM, ITERS = 5, 20


def timed(fn):
    for _ in range(WARM):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4096,262144,640000")
    args = ap.parse_args()
    try:
        import _raymarching as ref
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"unavailable": f"oracle/_ref/_raymarching.so not importable: {e!r}"[:300]}))
        return
    from stable_nerf_b200 import _lib, synthetic as syn
    lib = _lib.load()
    P, chk = _lib.ptr, _lib.check
    dev = torch.device("cuda", 0)
    S = _lib.stream()  # torch's current stream = the legacy default stream the reference launches on
    bound, C, H, max_steps, T_thresh, CH = 1.0, 1, 128, 1024, 1e-4, 3
    grid = syn.occupancy_grid(lego_like=True, seed=0)
    bitfield = torch.from_numpy(syn.pack_bitfield(grid)).to(dev)
    grid_t = torch.from_numpy(grid.astype(np.float32)).to(dev).contiguous()
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=dev)
    out = {"unit": "us per call", "iters": ITERS, "sizes": {}}

    for N in [int(s) for s in args.sizes.split(",")]:
        if N == 640000:
            ro, rd = syn.full_frame()
        else:
            ro, rd = syn.train_batch(N, seed=3)
        rays_o, rays_d = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        r = {}

        # ---- near_far_from_aabb
        r["near_far_from_aabb"] = {
            "reference": timed(lambda: ref.near_far_from_aabb(rays_o, rays_d, aabb, N, 0.2, nears, fars)),
            "ours": timed(lambda: chk(lib.snerf_near_far_from_aabb(P(rays_o), P(rays_d), P(aabb), N, 0.2, P(nears), P(fars), S), "nf"))}

        # ---- march_rays_train, steady state: M = the exact sample total rounded up to 128 (raymarching.py:199-203)
        noises = torch.zeros(N, device=dev)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        from stable_nerf_b200.raymarching import march_train_workspace_bytes
        ws_bytes = march_train_workspace_bytes(lib, N, max_steps)  # what the operator wrapper and TrainStep allocate
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        geom = (bound, 0.0, max_steps, N, C, H)
        chk(lib.snerf_march_rays_train_count(P(rays_o), P(rays_d), P(bitfield), *geom, P(nears), P(fars), P(counter), P(noises),
                                             P(ws), ws_bytes, S), "count")
        total = int(counter[0].item())
        M = (total + 127) // 128 * 128 + 128
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        n_samples = torch.empty(1, dtype=torch.int32, device=dev)
        bufs = {}

        def ref_march():
            # the wrapper's allocations (raymarching.py:205-207: torch.zeros) + counter.zero_() (nerf/renderer.py:97) + kernel
            bufs["x"] = torch.zeros(M, 3, device=dev)
            bufs["d"] = torch.zeros(M, 3, device=dev)
            bufs["dl"] = torch.zeros(M, 2, device=dev)
            counter.zero_()
            ref.march_rays_train(rays_o, rays_d, bitfield, bound, 0.0, max_steps, N, C, H, M, nears, fars, bufs["x"], bufs["d"],
                                 bufs["dl"], rays, counter, noises)
        xyzs, dirs, deltas = torch.empty(M, 3, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 2, device=dev)

        def our_march():
            counter.zero_()
            chk(lib.snerf_march_rays_train_count(P(rays_o), P(rays_d), P(bitfield), *geom, P(nears), P(fars), P(counter),
                                                 P(noises), P(ws), ws_bytes, S), "count")
            chk(lib.snerf_march_rays_train_write(P(rays_o), P(rays_d), P(bitfield), *geom, M, P(nears), P(fars), P(xyzs), P(dirs),
                                                 P(deltas), P(rays), P(noises), 1, P(n_samples), P(ws), ws_bytes, S), "write")
        r["march_rays_train"] = {"reference": timed(ref_march), "ours": timed(our_march), "samples": total, "M": M}
        if N <= 4096:  # the reference's first-epoch path: N*max_steps rows zero-filled, then the kernel
            Mq = N * max_steps

            def ref_march_q8():
                bufs["x"] = torch.zeros(Mq, 3, device=dev)
                bufs["d"] = torch.zeros(Mq, 3, device=dev)
                bufs["dl"] = torch.zeros(Mq, 2, device=dev)
                counter.zero_()
                ref.march_rays_train(rays_o, rays_d, bitfield, bound, 0.0, max_steps, N, C, H, Mq, nears, fars, bufs["x"],
                                     bufs["d"], bufs["dl"], rays, counter, noises)
            r["march_rays_train_first_epoch_path"] = {"reference": timed(ref_march_q8), "ours": r["march_rays_train"]["ours"],
                                                      "note": "reference: N*max_steps rows zero-filled per call (raymarching.py:196-207)"}
        bufs.clear()
        our_march()  # canonical packing for the compositing below
        torch.cuda.synchronize()

        # ---- composite_rays_train forward / backward on the packed samples
        g = torch.Generator(device="cpu").manual_seed(N)
        sig = (torch.rand(M, generator=g) ** 3 * 20.0).to(dev)
        rgb = torch.rand(M, CH, generator=g).to(dev)
        wsum, depth, image = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, CH, device=dev)
        r["composite_rays_train_forward"] = {
            "reference": timed(lambda: ref.composite_rays_train_forward(sig, rgb, deltas, rays, M, N, T_thresh, CH, wsum, depth, image)),
            "ours": timed(lambda: chk(lib.snerf_composite_rays_train_forward(P(sig), P(rgb), P(deltas), P(rays), M, N, T_thresh, CH,
                                                                            P(wsum), P(depth), P(image), S), "cf"))}
        g_ws, g_img = torch.rand(N, generator=g).to(dev), torch.rand(N, CH, generator=g).to(dev)
        gs, gr = torch.empty(M, device=dev), torch.empty(M, CH, device=dev)

        def ref_cbwd():
            zs, zr = torch.zeros_like(sig), torch.zeros_like(rgb)  # raymarching.py:283-284
            ref.composite_rays_train_backward(g_ws, g_img, sig, rgb, deltas, rays, wsum, image, M, N, T_thresh, CH, zs, zr)
        r["composite_rays_train_backward"] = {
            "reference": timed(ref_cbwd),
            "ours": timed(lambda: chk(lib.snerf_composite_rays_train_backward_ex(
                P(g_ws), P(g_img), P(sig), P(rgb), P(deltas), P(rays), P(wsum), P(image), M, N, T_thresh, CH, P(gs), P(gr),
                P(n_samples), S), "cb"))}

        # ---- inference: first loop iteration (all rays alive, n_step = 1, nerf/renderer.py:146) and its compositing
        n_step = 1
        Mi = (N * n_step + 127) // 128 * 128 + 128  # the reference's align always adds
        alive = torch.arange(N, dtype=torch.int32, device=dev)
        rays_t = nears.clone()
        ix, idr, idl = torch.empty(Mi, 3, device=dev), torch.empty(Mi, 3, device=dev), torch.empty(Mi, 2, device=dev)

        def ref_march_inf():
            bufs["x"] = torch.zeros(Mi, 3, device=dev)  # raymarching.py:335-337
            bufs["d"] = torch.zeros(Mi, 3, device=dev)
            bufs["dl"] = torch.zeros(Mi, 2, device=dev)
            nz = torch.zeros(N, device=dev)             # :339-342
            ref.march_rays(N, n_step, alive, rays_t, rays_o, rays_d, bound, 0.0, max_steps, C, H, bitfield, nears, fars,
                           bufs["x"], bufs["d"], bufs["dl"], nz)
        r["march_rays"] = {
            "reference": timed(ref_march_inf),
            "ours": timed(lambda: chk(lib.snerf_march_rays_ex(N, n_step, P(alive), P(rays_t), P(rays_o), P(rays_d), bound, 0.0,
                                                              max_steps, C, H, P(bitfield), P(nears), P(fars), P(ix), P(idr), P(idl),
                                                              None, Mi, S), "mr"))}
        isig = (torch.rand(Mi, generator=g) ** 3 * 20.0).to(dev)
        irgb = torch.rand(Mi, CH, generator=g).to(dev)
        iws, idep, iimg = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, CH, device=dev)
        alive0, t0 = alive.clone(), rays_t.clone()

        def reset():
            alive.copy_(alive0)
            rays_t.copy_(t0)
            iws.zero_(); idep.zero_(); iimg.zero_()
        t_reset = timed(reset)

        def ref_cinf():
            reset()
            ref.composite_rays(N, n_step, T_thresh, CH, alive, rays_t, isig, irgb, idl, iws, idep, iimg)

        def our_cinf():
            reset()
            chk(lib.snerf_composite_rays(N, n_step, T_thresh, CH, P(alive), P(rays_t), P(isig), P(irgb), P(idl), P(iws), P(idep),
                                         P(iimg), S), "ci")
        r["composite_rays"] = {"reference": max(timed(ref_cinf) - t_reset, 0.0), "ours": max(timed(our_cinf) - t_reset, 0.0),
                               "note": "state reset before each call, its own time subtracted"}
        out["sizes"][str(N)] = r

    # ---- packbits over one cascade (2 097 152 cells)
    bits = torch.empty(grid_t.numel() // 8, dtype=torch.uint8, device=dev)
    n8 = grid_t.numel() // 8
    out["packbits_2M_cells"] = {
        "reference": timed(lambda: ref.packbits(grid_t.view(1, -1), n8, 0.01, bits)),
        "ours": timed(lambda: chk(lib.snerf_packbits(P(grid_t), n8, 0.01, P(bits), S), "pb"))}
    for size in out["sizes"].values():
        for op in size.values():
            op["speedup"] = round(op["reference"] / op["ours"], 2) if op["ours"] > 0 else None
            op["reference"], op["ours"] = round(op["reference"], 2), round(op["ours"], 2)
    pb = out["packbits_2M_cells"]
    pb["speedup"] = round(pb["reference"] / pb["ours"], 2)
    print("REF_GPU_KERNELS " + json.dumps(out))


if __name__ == "__main__":
    main()
