#!/usr/bin/env python
"""Benchmark of the PDFNet depth-branch + fusion + MANO hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision bf16|fp32]

One "step" = one pass of the hot path over one batch of synthetic RGB-D frames:
depth -> per-hand clouds (device depth2pcl) -> 3-level pixel->point gather + SFT0 -> SA1 -> SFT1
-> SA2 -> SFT2 -> global MLP + max -> final SFT(1024) with centre features -> mano_head ->
Split_coeff -> MANO LBS.  Workload = BASELINE.json configs[2] restricted to the hot path
(the RGB ResNet-50 neck and the GCN decoder are out of scope; their
outputs — feature pyramid, hand masks, centre features, centre indices — are synthetic inputs).

Prints ONE JSON line (rank 0).  Multi-GPU: one process per GPU (torchrun), frames sharded,
no data-path collective ("scaling": "weak", fixed frames per GPU); time = max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rgbd_frames_per_sec_hot_path"
UNIT = "frames/s"

# algorithmic work per unit (SURVEY.md section 8d / BASELINE.md section 3)
FLOP_SA1, FLOP_SA2, FLOP_GLOBAL = 817889280, 1080033280, 235274240
FLOP_SFT1, FLOP_SFT2 = 25559040, 67502080


def make_opt(R):
    return types.SimpleNamespace(SAMPLE_NUM=1024, INPUT_FEATURE_NUM=3, knn_K=64, sample_num_level1=512,
                                 sample_num_level2=128, ball_radius=0.015, ball_radius2=0.04, default_resolution=R,
                                 PCA_SZ=63)


D2P_SEED = 317          # seed of the kernel-generated depth2pcl randomness (keys / permutation), reference seed opts.py:56


def make_inputs(n_frames, R, seed, pyramid="bf16-nhwc", masks="u8"):
    """Synthetic host inputs of the hot path for ``n_frames`` frames.

    pyramid: 'bf16-nhwc' = what an autocast, channels-last RGB neck emits (BASELINE cfg3 "forward bf16");
             'fp32-nchw' = the reference's own fp32 NCHW maps; 'fp32-nhwc' = fp32 channels-last.
    masks:   'u8' (hand / not hand) or 'f32' (the reference's dtype).
    The values are the same in every variant: fp32 maps hold the bf16-rounded numbers, so the CPU arm (fp32 NCHW)
    and every GPU variant see identical inputs."""
    from pdfnet_b200 import synth
    depth, mask, K, valid = synth.rgbd_frames(n_frames, R, seed=seed)
    emb = [e.bfloat16() for e in synth.pyramid(n_frames, R, seed=seed)]
    if pyramid == "bf16-nhwc":
        emb = [e.contiguous(memory_format=torch.channels_last) for e in emb]
    elif pyramid == "fp32-nhwc":
        emb = [e.float().contiguous(memory_format=torch.channels_last) for e in emb]
    else:
        emb = [e.float() for e in emb]
    if masks == "u8":
        mask = (mask > 0.5).to(torch.uint8)
    g = torch.Generator().manual_seed(seed)
    center = torch.randn((n_frames, 2, 1024), generator=g)
    ind = torch.randint(0, (R // 4) ** 2, (n_frames, 2), generator=g)
    Kinv = torch.from_numpy(np.stack([np.linalg.inv(k) for k in K.numpy()]))   # host, as utils.py:269
    return dict(depth=depth, mask=mask, K=K, Kinv=Kinv, valid=valid, l0=emb[0], l1=emb[1], l2=emb[2], center=center,
                ind=ind)


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path (the reference is
# Python/PyTorch and cannot travel to the GPU box; oracle/pdf_oracle.py restates it 1:1 and is
# pinned to it by tests/golden).  This is the only place bench.py executes oracle/.
# ----------------------------------------------------------------------------------------------

def cpu_hot_path(inp, R, mano_tables, state, dec=None, stage_times=None):
    from oracle import pdf_oracle as O
    opt = make_opt(R)
    B = inp["depth"].shape[0]

    def tick(name, t0):
        if stage_times is not None:
            stage_times[name] = stage_times.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    t = time.perf_counter()
    keys, perm = O.d2p_seeded_randomness(D2P_SEED, 2 * B, R * R)
    keys, perm = keys.reshape(B, 2, R * R), perm.reshape(B, 2, 1024)
    mask = inp["mask"].float().numpy()
    chooses, clouds = [], []
    for b in range(B):
        ch, cl = O.depth2pcl(inp["depth"][b].numpy(), mask[b:b + 1], inp["K"][b].numpy(),
                             inp["valid"][b:b + 1].numpy(), keys[b], perm[b])
        chooses.append(torch.from_numpy(ch))
        clouds.append(torch.from_numpy(cl))
    choose, cloud = torch.stack(chooses), torch.stack(clouds)
    t = tick("depth2pcl (a1,a2)", t)
    emb = [inp[k].float().contiguous() for k in ("l0", "l1", "l2")]
    with torch.no_grad():
        feats = [O.pointnet_plus_forward(state["pointnet"], cloud[:, h], emb, choose[:, h], opt) for h in (0, 1)]
        t = tick("PointNet_Plus x2 (a4-a9)", t)
        fuse = O.sft_layer(torch.cat(feats, 1).transpose(1, 2).contiguous(), inp["center"], state["sft"], "")
        t = tick("fusion SFT (a10)", t)
        th_l = O.mano_head(feats[0][:, 0], state["mano_head"])
        th_r = O.mano_head(feats[1][:, 0], state["mano_head"])
        sl = O.split_coeff(th_l, inp["ind"][:, 0], inp["K"], R, 4)
        sr = O.split_coeff(th_r, inp["ind"][:, 1], inp["K"], R, 4)
        vl, jl = O.mano_lbs(mano_tables["left"], sl[0], sl[1], sl[2], side="left")
        vr, jr = O.mano_lbs(mano_tables["right"], sr[4], sr[5], sr[6], side="right")
        t = tick("mano_head + Split_coeff + LBS (a11-a14)", t)
        out = [fuse, torch.stack((vl, vr), 1), torch.stack((jl, jr), 1)]
        if dec is not None:
            sd_d, assets = dec
            res = O.gcn_decoder_forward(sd_d, assets, fuse)
            out += [res["verts3d_left"], res["verts3d_right"]]
            out += [O.regress_joints(O.full_regressor(mano_tables[s]["J_regressor"]), res["verts3d_" + s])
                    for s in ("left", "right")]
            t = tick("GCN decoder + full_regressor (f3,a15)", t)
    return out


def load_states():
    from pdfnet_b200 import synth
    return dict(pointnet=synth.pointnet_plus_state(317), sft=synth.fusion_sft_state(317),
                mano_head=synth.mano_head_state(317, std=0.05))


def load_mano_tables():
    """Real MANO tables exported to tests/golden (npz); synthetic MANO-shaped tables otherwise."""
    from pdfnet_b200 import synth
    out = {}
    for side in ("left", "right"):
        p = os.path.join(ROOT, "tests", "golden", "mano_%s.npz" % side)
        out[side] = dict(np.load(p)) if os.path.exists(p) else synth.to_numpy(synth.synthetic_mano_tables())
    return out


def load_decoder_state():
    from pdfnet_b200 import synth
    assets = dict(np.load(os.path.join(ROOT, "tests", "golden", "gcn_assets.npz")))
    return synth.decoder_state(317, assets["upsample"]), assets


def time_cpu(sample_frames, R, steps, warmup, with_decoder, stage_times=None):
    torch.set_num_threads(os.cpu_count() or 1)
    inp = make_inputs(sample_frames, R, seed=317, pyramid="fp32-nchw", masks="u8")
    tables, state = load_mano_tables(), load_states()
    dec = load_decoder_state() if with_decoder else None
    for _ in range(warmup):
        cpu_hot_path(inp, R, tables, state, dec)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_hot_path(inp, R, tables, state, dec, stage_times)
        ts.append(time.perf_counter() - t0)
    return ts


def config_dict(args, world, B, extra=None):
    """The keys both arms print (the driver compares them): workload, frames per GPU and step, resolution, precision,
    input formats, parallelism."""
    c = {"workload": workload_name(args), "frames_per_gpu": B, "resolution": args.res, "precision": args.precision,
         "parallelism": "dp%d" % world, "pyramid": args.pyramid, "masks": args.masks, "with_decoder": args.with_decoder,
         "randomness": "depth2pcl subset keys / permutation generated from seed %d (counter-based; the reference "
                       "draws them from np.random)" % D2P_SEED}
    c.update(extra or {})
    return c


def run_reference(args):
    """Reference arm: the oracle port of the reference's CPU path on all host threads, on a bounded sample of
    the same workload (same inputs, same stages incl. the GCN decoder), rank 0 only."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sample = args.cpu_sample_frames
    stage_t = {}
    ts = time_cpu(sample, args.res, args.steps, args.warmup, args.with_decoder, stage_t)
    ms = 1e3 * sum(ts) / len(ts)
    val = sample / (ms / 1e3)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, args.gpus, args.frames, {"sample_frames_per_step": sample}),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d frames/step of the same workload (of %d per GPU step), torch CPU fp32, %d threads"
                                   % (sample, args.frames, cores),
                         "stage_ms_per_frame": {k: round(1e3 * v / (len(ts) * sample), 3) for k, v in stage_t.items()}},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------

def workload_name(args):
    return ("cfg3-hotpath: %d frames/GPU x 2 hands, %dx%d depth, 1024-pt clouds, N1=512 N2=128 K=64 r2=(0.015,0.04); "
            "depth2pcl+pyramid gather/SFT+SA1+SA2+global MLP+fusion SFT+mano_head+Split_coeff+LBS%s"
            % (args.frames, args.res, args.res,
               "+GCN decoder+full_regressor" if getattr(args, "with_decoder", False) else ""))


class ClockSampler(object):
    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# algorithmic bytes per unit of the HBM-bound stages
def stage_bytes(args, n_frames):
    R = args.res
    mask_b = 1 if args.masks == "u8" else 4
    feat_b = 2 if args.pyramid.startswith("bf16") else 4
    d2p = R * R * 4 + 2 * R * R * mask_b + 2 * 1024 * (12 + 8)                       # per frame
    pyr_read = 1024 * 3 * feat_b + 512 * 64 * feat_b + 128 * 256 * feat_b + 1024 * 8 + 1024 * 12
    pyr_write = 1024 * 12 + 512 * 64 * feat_b + 128 * 256 * feat_b                  # pts0 + condition rows / images
    return {"depth2pcl": d2p * n_frames, "pyramid_gather": (pyr_read + pyr_write) * 2 * n_frames,
            "knn1": 143360 * 2 * n_frames, "knn2": 38912 * 2 * n_frames}


def run_ours(args):
    from pdfnet_b200 import parallel
    torch.set_grad_enabled(False)         # inference workload: fused kernels (a graph would select training.py)
    rank, world, local = parallel.init_distributed("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from pdfnet_b200 import HandFusion, ManoLayer, _lib, mano_tail_pair, ops, profiling
    from pdfnet_b200.graph import CapturedStep

    R, B = args.res, args.frames                      # frames per GPU (weak scaling)
    opt = make_opt(R)
    host = make_inputs(B, R, seed=317 + rank, pyramid=args.pyramid, masks=args.masks)
    # staging buffers the CPU only writes: write-combined page-locked memory (not snooped while the copy engines
    # read it; matters when several ranks pull from the same host memory); --host-alloc pinned = torch's pin_memory
    pinned = {k: parallel.pinned_like(v, write_combined=(args.host_alloc == "wc")) for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    state = load_states()
    tables = load_mano_tables()
    model = HandFusion(opt, precision=args.precision)
    sd = {"pointnet_plus." + k: v for k, v in state["pointnet"].items()}
    sd.update({"sft." + k: v for k, v in state["sft"].items()})
    sd.update(state["mano_head"])
    model.load_state_dict(sd, strict=False)
    model = model.to(dev).eval()
    mano_l = ManoLayer(tables["left"], center_idx=None).to(dev)
    mano_r = ManoLayer(tables["right"], center_idx=None).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    dec = None
    if args.with_decoder:                             # SURVEY 8f row f3 + a15: GCN decoder consuming fuse_feat, then the joints
        from pdfnet_b200.decoder import decoder
        sd_d, assets = load_decoder_state()
        dec = decoder(assets, precision="fp32" if args.precision == "fp32" else "bf16x3")
        dec.load_state_dict(sd_d)
        dec.set_joint_regressors(tables["left"]["J_regressor"], tables["right"]["J_regressor"])
        dec = dec.to(dev).eval()

    mano_side = torch.cuda.Stream(device=dev)

    def hot_stage_a(d):
        """Stage A of the split pass: cloud builder + pixel -> point gather (the only kernels that read the host-resident
        pyramid in zero-copy mode) -> (cloud, pts0, cond1 image, cond2 image)."""
        choose, cloud, _ = ops.depth2pcl(d["depth"], d["mask"], d["Kinv"], d["valid"], seed=D2P_SEED)
        return (cloud,) + tuple(model.gather(cloud, [d["l0"], d["l1"], d["l2"]], choose))

    def hot_path(d, overlap=True, pre=None):
        """One pass.  overlap: the MANO branch (head, Split_coeff, LBS - it only needs the un-fused features) runs
        on a side stream beside the fusion SFT and the GCN decoder (a fork / join inside the captured graph);
        overlap=False keeps everything on one stream for the per-stage timings.  pre: the outputs of hot_stage_a
        (stage B of the split pass: everything after the gather)."""
        side = mano_side if (overlap and dec is not None) else None
        if pre is not None:
            fused, theta = model(pre[0], None, None, d["center"], with_mano=True, mano_stream=side, gathered=tuple(pre[1:]))
        else:
            with profiling.stage("depth2pcl"):
                choose, cloud, _ = ops.depth2pcl(d["depth"], d["mask"], d["Kinv"], d["valid"], seed=D2P_SEED)
            fused, theta = model(cloud, [d["l0"], d["l1"], d["l2"]], choose, d["center"], with_mano=True, mano_stream=side)
        if side is not None:
            with torch.cuda.stream(side):
                verts, joints, _ = mano_tail_pair(theta, d["ind"], d["K"], mano_l, mano_r, input_res=R)
        else:
            with profiling.stage("mano_tail"):
                verts, joints, _ = mano_tail_pair(theta, d["ind"], d["K"], mano_l, mano_r, input_res=R)
        if dec is not None:
            with profiling.stage("gcn_decoder"):
                result, _, _, other = dec(fused[:, 0], fused[:, 1], None)
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)         # join
                for t in (verts, joints):
                    t.record_stream(torch.cuda.current_stream())
            return (fused, verts, joints, result["verts3d"]["left"], result["verts3d"]["right"],
                    other["joints3d"]["left"], other["joints3d"]["right"])
        return fused, verts, joints

    def hot_front(d):
        """Pipeline stage 1 (everything up to fuse_feat + the MANO branch, which joins before the stage ends)."""
        choose, cloud, _ = ops.depth2pcl(d["depth"], d["mask"], d["Kinv"], d["valid"], seed=D2P_SEED)
        fused, theta = model(cloud, [d["l0"], d["l1"], d["l2"]], choose, d["center"], with_mano=True, mano_stream=mano_side)
        with torch.cuda.stream(mano_side):
            verts, joints, _ = mano_tail_pair(theta, d["ind"], d["K"], mano_l, mano_r, input_res=R)
        torch.cuda.current_stream().wait_stream(mano_side)
        for t in (verts, joints):
            t.record_stream(torch.cuda.current_stream())
        return fused, (fused, verts, joints)

    def hot_back(fused):
        """Pipeline stage 2: GCN decoder + joint regressor on a batch's fuse_feat."""
        result, _, _, other = dec(fused[:, 0], fused[:, 1], None)
        return (result["verts3d"]["left"], result["verts3d"]["right"], other["joints3d"]["left"], other["joints3d"]["right"])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, steps, warmup, min_seconds=0.0):
        """CUDA events around every step, L2 flushed in between (outside the events); -> mean ms (max over ranks)."""
        for _ in range(warmup):
            step_fn()
        barrier()
        evs, t0 = [], time.perf_counter()
        while len(evs) < steps or (time.perf_counter() - t0) < min_seconds:
            flush.fill_(1)                            # evict L2 between timed iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
            if min_seconds and len(evs) % 64 == 0:
                torch.cuda.synchronize()              # keep the launch queue bounded in the long loop
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        return parallel.max_over_ranks(ms, dev), len(evs)

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms_eager, _ = timed_loop(lambda: hot_path(resident), args.steps, args.warmup)
    launches = (_lib.launch_count() - l0) // (args.steps + args.warmup)
    ms_dev, graphed, step = ms_eager, False, None
    if not args.no_graph:
        # the same kernels replayed as one CUDA graph (no Python / ctypes launch cost in the step)
        try:
            step = CapturedStep(lambda: hot_path(resident))
            launches = step.launches                  # exact: the library kernels inside the captured step
            (ms_dev, _), graphed = timed_loop(step.replay, args.steps, args.warmup), True
        except Exception as e:                        # keep the eager measurement rather than lose the line
            sys.stderr.write("bench: CUDA-graph capture failed (%s); reporting the eager launch path\n" % e)
            torch.cuda.synchronize()
            args.no_graph, step = True, None
    # ---- the same work software-pipelined over consecutive batches: replay i = point branch of batch i beside the
    # GCN decoder of batch i-1 (pdfnet_b200.graph.PipelinedStep); checked against the serial step's outputs ----
    pipelined = None
    if step is not None and dec is not None and args.pipeline:
        try:
            from pdfnet_b200.graph import PipelinedStep
            prio = int(os.environ.get("PDF_PIPE_BACK_PRIO", "0"))
            if prio:
                dec._side = torch.cuda.Stream(device=dev, priority=prio)
            pstep = PipelinedStep(lambda: hot_front(resident), hot_back, back_priority=prio)
            ms_pipe, _ = timed_loop(pstep.replay, args.steps, args.warmup)
            fo, bo = pstep.replay()
            torch.cuda.synchronize()
            ref_out = step.replay()
            torch.cuda.synchronize()
            same = all(torch.equal(a, b) for a, b in zip(tuple(fo) + tuple(bo), ref_out))
            dr = pstep.flush()
            torch.cuda.synchronize()
            same = same and all(torch.equal(a, b) for a, b in zip(dr, ref_out[3:]))
            pipelined = {"ms_per_step": ms_pipe, "value": world * B / (ms_pipe * 1e-3), "unit": UNIT,
                         "launches": pstep.launches, "outputs_equal_serial_step": bool(same), "back_priority": prio}
        except Exception as e:
            pipelined = {"error": str(e)[:200]}
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    # ---- sustained: the same step back to back for >= args.sustained_seconds (power-capped clocks) ----
    sustained = None
    if args.sustained_seconds > 0:
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        fn = step.replay if step is not None else (lambda: hot_path(resident))
        ms_sus, n_sus = timed_loop(fn, args.steps, 0, min_seconds=args.sustained_seconds)
        sustained = {"value": world * B / (ms_sus * 1e-3), "unit": UNIT, "ms_per_step": ms_sus, "steps": n_sus,
                     "seconds": args.sustained_seconds, "clocks": s2.stop() if rank == 0 else None}

    # ---- end to end through the public API with HOST buffers ----
    # Every step takes ALL hot-path inputs from page-locked host memory and returns the results there.  The batch is
    # cut into chunks and the staging buffers are double-buffered ACROSS steps: while chunk c of step i computes, the
    # copy stream already moves later chunks / the next step's inputs, and a third stream returns the outputs.
    # Two hand-offs of the RGB feature pyramid (92 % of the input bytes) are measured:
    #   zero-copy: the bf16 channels-last maps STAY in page-locked host memory and the gather kernel reads the pixels
    #              `choose` selects in place over the PCIe link (2 x (1024 x 6 + 512 x 128 + 128 x 512) B of a frame's
    #              4.6 MB); depth, masks, intrinsics, centre features / indices are copied as before;
    #   copy:      every input, the whole pyramid included, is copied to device staging buffers first.
    input_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    n_chunks = max(1, min(args.e2e_chunks, B))
    bounds_chunks = [parallel.shard_range(B, c, n_chunks) for c in range(n_chunks)]
    copy_stream, out_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    names = ("fused", "verts", "joints", "gcn_verts_left", "gcn_verts_right", "gcn_joints_left", "gcn_joints_right")
    PYR = ("l0", "l1", "l2")
    zc_possible = args.pyramid == "bf16-nhwc" and args.precision == "bf16"
    d2h_bytes = [0]

    def pcie_rx_sampler(stop, rows):
        """NVML PCIe receive throughput of this GPU (KB/s over 20 ms windows) while the e2e loop runs."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            while not stop.is_set():
                rows.append(pynvml.nvmlDeviceGetPcieThroughput(h, pynvml.NVML_PCIE_UTIL_RX_BYTES))
        except Exception:
            pass

    zc_arg = args.e2e_zero_copy_levels
    if zc_arg == "auto":
        zc_arg = "l1,l2" if world <= 2 else "l0,l1,l2"      # up to 2 ranks the host feeds every link at full speed (measured)
    zc_levels = tuple(k for k in PYR if k in zc_arg.split(","))
    comp_streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, args.e2e_streams))]
    pinned_alt = {}
    zc_px_bytes = {"l0": 1024 * 6, "l1": 512 * 128, "l2": 128 * 512}          # per cloud (SURVEY 8d, bf16 features)

    def measure_e2e(zero_copy, prefetch=False):
        """prefetch: the pass is split at the gather.  Stage A (cloud builder + gather, the part that waits on the PCIe
        link in zero-copy mode) of step i+1 runs on one stream while stage B (everything else, the WHOLE batch as one
        graph: the decoder's launch chain is paid once per step, not once per chunk) of step i runs on another."""
        zc = zc_levels if zero_copy else ()
        bounds = [(0, B)] if prefetch else bounds_chunks
        n_sets = 3 if prefetch else 2                 # split schedule: a third staging set keeps the copies off stage B's heels
        # host side: two page-locked pyramid sets read in place alternately (no step re-reads the addresses of the
        # step before it); device side: double-buffered staging for everything that is copied
        # set-up (page-locked allocations, graph captures) can fail on ONE rank only; the timed part below holds
        # barriers, so the ranks first agree that every one of them got through it
        fail = None
        chunk_steps = a_steps = staging = probe = out_host = None
        try:
            for k in zc:                                  # second page-locked copy of the maps read in place (kept across modes)
                if k not in pinned_alt:
                    pinned_alt[k] = parallel.pinned_like(pinned[k], write_combined=(args.host_alloc == "wc"))
            host_sets = [pinned, {k: (pinned_alt[k] if k in zc else pinned[k]) for k in pinned}]
            staging = [{k: (host_sets[st % 2][k] if k in zc else torch.empty_like(resident[k])) for k in pinned}
                       for st in range(n_sets)]
            copied = [k for k in pinned if k not in zc]
            copy_bytes = sum(pinned[k].numel() * pinned[k].element_size() for k in copied)
            zc_bytes = 2 * B * sum(zc_px_bytes[k] for k in zc)                            # algorithmic bytes read in place
            chunk_steps = a_steps = None
            if prefetch:                                  # two graphs per staging set: stage A, stage B on A's outputs
                a_steps = [CapturedStep(lambda st=st: hot_stage_a(st), warmup=2) for st in staging]
                chunk_steps = [[CapturedStep(lambda st=st, a=a: hot_path(st, pre=a.outputs), warmup=2)]
                               for st, a in zip(staging, a_steps)]
            elif not args.no_graph:                       # one captured graph per (staging set, chunk)
                try:
                    chunk_steps = [[CapturedStep(lambda lo=lo, hi=hi, st=st: hot_path({k: v[lo:hi] for k, v in st.items()}),
                                                 warmup=2) for lo, hi in bounds] for st in staging]
                except Exception as e:
                    sys.stderr.write("bench: CUDA-graph capture of the e2e chunks failed (%s); eager chunks\n" % e)
                    torch.cuda.synchronize()
                    chunk_steps = None
            probe = hot_path({k: v[bounds[0][0]:bounds[0][1]] for k, v in staging[0].items()})
            out_host = [{n: torch.empty((B,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory() for n, t in zip(names, probe)}
                        for _ in range(n_sets)]
        except Exception as e:
            fail = str(e)[:200]
            torch.cuda.synchronize()
        if parallel.max_over_ranks(1.0 if fail else 0.0, dev) > 0:
            del chunk_steps, a_steps, staging, probe, out_host
            torch.cuda.empty_cache()
            raise RuntimeError("e2e set-up failed on %s: %s" % ("this rank" if fail else "another rank", fail))
        d2h_bytes[0] = sum(v.numel() * v.element_size() for v in out_host[0].values())
        consumed = [None] * n_sets                    # event: the compute of the step that last used this set is done
        drained = [None] * n_sets                     # event: its outputs have left the device

        def e2e_step(i):
            # chunks alternate between the compute streams: while one chunk computes, the next one's depth2pcl and
            # (in zero-copy mode) its gather over the PCIe link are already under way
            st = i % n_sets
            for ev in consumed[st] or ():
                copy_stream.wait_event(ev)
            events = []
            with torch.cuda.stream(copy_stream):
                for lo, hi in bounds:
                    for k in copied:
                        staging[st][k][lo:hi].copy_(pinned[k][lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    events.append(ev)
            for ci, ((lo, hi), ev) in enumerate(zip(bounds, events)):
                S = comp_streams[(i * len(bounds) + ci) % len(comp_streams)]
                if prefetch:                          # stage A on stream 0 (ordered after this set's last stage B through
                    S = comp_streams[1]               # consumed[st] -> copy stream -> ev), stage B on stream 1
                    with torch.cuda.stream(comp_streams[0]):
                        comp_streams[0].wait_event(ev)
                        a_steps[st].replay()
                        ev = torch.cuda.Event()
                        ev.record(comp_streams[0])
                with torch.cuda.stream(S):
                    if drained[st] is not None:
                        S.wait_event(drained[st])     # graph outputs of this set are free to be overwritten
                    S.wait_event(ev)
                    outs = chunk_steps[st][ci].replay() if chunk_steps else \
                        hot_path({k: v[lo:hi] for k, v in staging[st].items()})
                    done = torch.cuda.Event()
                    done.record(S)
                with torch.cuda.stream(out_stream):
                    out_stream.wait_event(done)
                    for n, t in zip(names, outs):
                        out_host[st][n][lo:hi].copy_(t, non_blocking=True)
                        t.record_stream(out_stream)
            consumed[st] = []
            for S in comp_streams:
                ev = torch.cuda.Event()
                ev.record(S)
                consumed[st].append(ev)
            drained[st] = torch.cuda.Event()
            drained[st].record(out_stream)

        for S in comp_streams:
            S.wait_stream(torch.cuda.current_stream())
        for i in range(args.warmup):
            e2e_step(i)
        barrier()
        stop, rows = threading.Event(), []
        th = threading.Thread(target=pcie_rx_sampler, args=(stop, rows), daemon=True)
        steps = max(args.steps, 40)                   # >= 40 steps: the region spans several 20 ms NVML windows
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        th.start()
        a.record()
        for S in comp_streams + [copy_stream]:
            S.wait_event(a)                           # nothing of the timed steps starts before the clock does
        for i in range(steps):
            e2e_step(args.warmup + i)
        main = torch.cuda.current_stream()
        for S in comp_streams:
            main.wait_stream(S)
        main.wait_stream(copy_stream)
        main.wait_stream(out_stream)                  # the last results are on the host when the clock stops
        b.record()
        torch.cuda.synchronize()
        stop.set()
        barrier()
        ms = parallel.max_over_ranks(a.elapsed_time(b) / steps, dev)
        # the zero-copy results are the copy path's results: same kernels on the same values
        nm = names[:len(probe)]                       # without the decoder the pass returns the first three only
        res = {n: out_host[(args.warmup + steps - 1) % n_sets][n].clone() for n in nm}
        th.join(timeout=1.0)
        rx = sorted(rows[1:-1]) if len(rows) > 2 else sorted(rows)
        rec = {"value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
               "h2d_bytes_per_step": copy_bytes + zc_bytes, "copied_bytes_per_step": copy_bytes,
               "zero_copy_bytes_per_step": zc_bytes, "d2h_bytes_per_step": d2h_bytes[0],
               "pcie_rx_gbs_nvml": round(rx[len(rx) // 2] * 1024 / 1e9, 2) if rx else None}
        rec["chunks"], rec["split_at_gather"] = len(bounds), bool(prefetch)
        # check: what arrived on the host equals the device-resident pass over the same chunks (kernel selection depends
        # on the chunk size, e.g. the decoder's 128-row gf layer, so each schedule is compared at its own chunking)
        torch.cuda.synchronize()
        ref = [hot_path({k: v[lo:hi] for k, v in resident.items()}) for lo, hi in bounds]
        torch.cuda.synchronize()
        diffs = [float((res[n].float() - torch.cat([r[j] for r in ref]).cpu().float()).abs().max()) for j, n in enumerate(nm)]
        rec["equals_device_resident_pass"] = bool(all(torch.equal(res[n], torch.cat([r[j] for r in ref]).cpu())
                                                      for j, n in enumerate(nm)))
        rec["max_abs_diff_vs_device_resident_pass"] = max(diffs)
        del ref
        del chunk_steps, a_steps, staging, probe, out_host
        torch.cuda.empty_cache()
        return rec, res

    e2e_modes = {}
    res_by_mode = {}
    split_ok = zc_possible and not args.no_graph and len(comp_streams) >= 2 and dec is not None
    for mode in ((("zero-copy-split",) if split_ok else ()) + ("zero-copy", "copy") if zc_possible else ("copy",)):
        try:
            e2e_modes[mode], res_by_mode[mode] = measure_e2e(mode != "copy", prefetch=(mode == "zero-copy-split"))
        except Exception as e:
            if mode == "copy":
                raise
            sys.stderr.write("bench: %s e2e failed (%s)\n" % (mode, e))
            torch.cuda.synchronize()
    if "zero-copy" in e2e_modes:                      # same chunking as the copy mode: bit-identical results
        e2e_modes["zero-copy"]["equals_copy_mode"] = bool(all(torch.equal(v, res_by_mode["copy"][n])
                                                              for n, v in res_by_mode["zero-copy"].items()))
    e2e_mode = args.e2e_pyramid
    if e2e_mode == "zero-copy":                       # the faster of the two zero-copy schedules measured in this run
        cands = [m for m in ("zero-copy-split", "zero-copy") if m in e2e_modes and e2e_modes[m]["equals_device_resident_pass"]]
        e2e_mode = min(cands, key=lambda m: e2e_modes[m]["ms_per_step"]) if cands else "copy"
    if e2e_mode not in e2e_modes:
        e2e_mode = "copy"
    ms_e2e = e2e_modes[e2e_mode]["ms_per_step"]
    h2d_bytes = e2e_modes[e2e_mode]["h2d_bytes_per_step"]
    d2h_bytes = d2h_bytes[0]
    del res_by_mode
    torch.cuda.synchronize()

    # ---- per-stage timing (roofline of the dominant kernel), same inputs, L2 flushed ----
    # The eager launch path is CPU-bound (one ctypes call per kernel): without a head start the event pairs would
    # bracket GPU idle time while Python prepares the next launch.  A device-side spin in front of every pass lets
    # the host queue the whole pass first, so each stage's events measure the kernels alone (cold L2, back to back).
    spin_cycles = int((0.045 if dec is not None else 0.012) * 1.9e9)
    for _ in range(2):
        hot_path(resident, overlap=False)
    profiling.enable(True)
    for _ in range(max(5, min(args.steps, 11))):
        flush.fill_(1)
        torch.cuda._sleep(spin_cycles)
        hot_path(resident, overlap=False)
    stages = profiling.summary()
    profiling.enable(False)
    if dec is not None and "gcn_decoder" in stages and not args.no_graph:
        # the eager decoder stage is launch-bound (~190 ctypes calls); inside the step it runs as part of the graph:
        # time the decoder alone as its own graph replay and report that (the eager figure is kept beside it)
        try:
            fz = hot_path(resident, overlap=False)[0].clone()
            dstep = CapturedStep(lambda: dec(fz[:, 0], fz[:, 1], None))
            ts = []
            for _ in range(max(5, min(args.steps, 11))):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); dstep.replay(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            stages["gcn_decoder_eager"] = stages["gcn_decoder"]
            stages["gcn_decoder"] = (len(ts), ts[len(ts) // 2])
            del dstep
        except Exception as e:
            sys.stderr.write("bench: decoder stage graph timing failed (%s)\n" % e)
            torch.cuda.synchronize()

    # ---- BASELINE cfg2 (SA microbench) and cfg5 (training step) sub-records, same process ----
    kernels = run_cfg2(args, dev=dev, emit=False)["kernels"] if (not args.no_sub and rank == 0) else None
    torch.cuda.empty_cache()

    # ---- BASELINE cfg4 AS STATED: a FIXED batch of 1024 frames sharded over the N GPUs (strong scaling, no collective);
    # at N = 8 that is the weak-scaling line above (128 frames per GPU), at N = 2 / 4 each GPU takes 512 / 256 frames ----
    cfg4 = None
    if not args.no_sub and world > 1 and 1024 % world == 0 and B == 128:
        F4 = 1024 // world
        try:
            if F4 == B:
                ms4 = ms_dev
            else:
                res4 = {k: v.to(dev) for k, v in make_inputs(F4, R, seed=4317 + rank, pyramid=args.pyramid, masks=args.masks).items()}
                step4 = CapturedStep(lambda: hot_path(res4))
                ms4, _ = timed_loop(step4.replay, min(args.steps, 5), 2)
                del step4, res4
                torch.cuda.empty_cache()
            cfg4 = {"workload": "cfg4: 1024 frames sharded over %d GPUs (%d per GPU), same step as the main line" % (world, F4),
                    "total_frames": 1024, "frames_per_gpu": F4, "ms_per_step": ms4, "value": 1024 / (ms4 * 1e-3),
                    "unit": UNIT, "scaling": "strong"}
        except Exception as e:
            cfg4 = {"error": str(e)[:200]}
            torch.cuda.synchronize()
    train = None
    if not args.no_sub:
        targs = argparse.Namespace(**vars(args))
        # BASELINE cfg5 is "training step bf16": bf16 operands (fp32 accumulation / activations) for the forward, dX
        # and dW GEMMs; the fp32-accurate mode (split-bf16 GEMMs, the library default) is timed beside it
        targs.frames, targs.precision, targs.steps, targs.warmup = 64, "bf16", min(args.steps, 5), 3
        try:
            train = run_cfg5(targs, emit=False, dist_ready=(rank, world, local))
            targs.precision, targs.steps = "fp32", min(args.steps, 3)
            alt = run_cfg5(targs, emit=False, dist_ready=(rank, world, local))
            if train is not None and alt is not None:
                train["fp32_accurate_mode"] = {k: alt[k] for k in ("value", "ms_per_step", "dtype", "final_loss") if k in alt}
        except Exception as e:                        # never lose the main line over the sub-record
            train = {"error": str(e)[:200]}
        torch.set_grad_enabled(False)

    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return

    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        j = json.load(open(pk))
        peaks = {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                 "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    n_clouds = 2 * B
    stage_ms = {k: t for k, (c, t) in stages.items()}
    flops = {"sa1": FLOP_SA1 * n_clouds, "sa2": FLOP_SA2 * n_clouds, "global_mlp": FLOP_GLOBAL * n_clouds,
             "sft1": FLOP_SFT1 * n_clouds, "sft2": FLOP_SFT2 * n_clouds}
    dom = max(flops, key=lambda k: stage_ms.get(k, 0.0))
    ach = flops[dom] / (stage_ms[dom] * 1e-3) / 1e12
    tensor_stage = args.precision == "bf16" and dom in ("sa1", "sa2", "global_mlp")
    # DRAM bytes per launch of that kernel: NOT measured in this run - read from the newest committed
    # `ncu --set full` capture (profiles/*_dram_traffic.json) and labelled as such
    traffic, traffic_source = None, None
    prof_dir = os.path.join(ROOT, "profiles")
    for cand in sorted((f for f in os.listdir(prof_dir) if f.endswith("_dram_traffic.json")), reverse=True) \
            if os.path.isdir(prof_dir) else []:
        tj = json.load(open(os.path.join(prof_dir, cand)))
        if dom in tj and args.precision == "bf16" and B == 128:
            traffic, traffic_source = tj[dom], "profiles/" + cand + " (committed ncu capture, not this run)"
        break
    roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_source,
                "peak_source": peaks["source"] + " burst bf16 (kernel timed alone with CUDA events)",
                "pipe": "tcgen05 bf16" if tensor_stage else "FFMA fp32 (stage not yet on tensor cores)",
                "launch_ms": stage_ms[dom]}
    stage_report = {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}
    stage_tflops = {k: round(flops[k] / (stage_ms[k] * 1e-3) / 1e12, 2) for k in flops if k in stage_ms}
    sb = stage_bytes(args, B)
    stage_hbm = {k: {"ms": round(stage_ms[k], 4), "algorithmic_gbs": round(sb[k] / (stage_ms[k] * 1e-3) / 1e9, 1),
                     "frac_of_hbm": round(sb[k] / (stage_ms[k] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
                 for k in sb if k in stage_ms}
    hot_flops = sum(flops.values()) + 16777216 * B
    if sustained is not None:
        sustained["hot_path_tflops"] = round(hot_flops / (sustained["ms_per_step"] * 1e-3) / 1e12, 1)
        sustained["frac_of_sustained_bf16"] = round(sustained["hot_path_tflops"] / peaks["bf16_tflops_sustained"], 4)
        sustained["note"] = "point-branch FLOPs (569.9 GFLOP at 128 frames) over the WHOLE step time, against the " \
                            "sustained cuBLAS figure; the step also holds the bandwidth / latency-bound stages"

    cores = os.cpu_count() or 1
    cpu_baseline = None
    if world == 1:
        st_t = {}
        cpu_ts = time_cpu(args.cpu_sample_frames, R, 3, 1, args.with_decoder, st_t)
        cpu_baseline = {"value": args.cpu_sample_frames / min(cpu_ts), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d frames of the same workload, oracle port (torch CPU fp32, %d threads), best of 3"
                                  % (args.cpu_sample_frames, cores),
                        "stage_ms_per_frame": {k: round(1e3 * v / (3 * args.cpu_sample_frames), 3) for k, v in st_t.items()}}

    line = {
        "metric": METRIC, "value": world * B / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": config_dict(args, world, B, {
            "l2": "256 MiB flush write between timed iterations; inputs %.2f GB > L2" % (input_bytes / 1e9),
            "launch": "one CUDA-graph replay per step" if graphed else "eager (one ctypes call per kernel)"}),
        "e2e": dict(e2e_modes[e2e_mode], **{
            "compute_streams": len(comp_streams), "host_alloc": args.host_alloc,
            "pyramid_handoff": e2e_mode, "zero_copy_levels": list(zc_levels) if e2e_mode != "copy" else [],
            "input_bytes_on_host_per_step": input_bytes,
            "h2d_gbs_per_rank": round(h2d_bytes / (ms_e2e * 1e-3) / 1e9, 2),
            "note": ("all hot-path inputs live in page-locked host memory every step.  depth, uint8 masks, K, centre "
                     "features and centre indices are copied to the device; the bf16 channels-last feature pyramid "
                     + ("is NOT copied: the gather kernel reads the pixels `choose` selects in place over the PCIe link "
                        "(zero_copy_bytes_per_step = algorithmic bytes of those pixels; two host pyramid sets read "
                        "alternately; split_at_gather: the cloud builder + gather of step i+1 run on one stream while "
                        "everything after the gather of step i runs on another, else the batch is cut into chunks that "
                        "alternate between the compute streams); " if e2e_mode != "copy" else "is copied whole; ")
                     + "staging double-buffered across steps (copy / compute / result streams); fused features, MANO "
                       "and GCN meshes and joints copied back to pinned host memory inside the timed region; "
                       "pcie_rx_gbs_nvml = the GPU's own PCIe receive counter (median 20 ms window) during the loop")}),
        "e2e_other_handoff": {k: v for k, v in e2e_modes.items() if k != e2e_mode},
        "gpu_launches": int(launches), "eager_ms_per_step": ms_eager, "clocks": clocks, "roofline": roofline,
        "stages_ms": stage_report, "stages_tflops": stage_tflops, "stages_hbm": stage_hbm,
        "pipelined": pipelined, "value_sustained": sustained, "kernels": kernels, "train": train, "cfg4": cfg4,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def run_cfg2(args, dev=None, emit=True):
    """BASELINE.json configs[1]: PointNet++ set-abstraction microbench, 64 clouds x 1024 points:
    FPS (1024 -> 512, cloud re-ordered so the FPS picks come first, interhand.py:857-900) ->
    ball query (r = 0.1 => r2 = 0.01, k = 64) -> fused point-MLP 3->64->64->128 + max (tcgen05).
    Reports per-kernel time, achieved algorithmic HBM GB/s (SURVEY 8d bytes) and TFLOP/s."""
    from pdfnet_b200 import PointNet_Plus, ops, synth
    torch.set_grad_enabled(False)
    if dev is None:
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
    B, N, N1, K = args.clouds, 1024, 512, 64
    pts = synth.clouds(B, seed=317).to(dev)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(317)).to(dev)
    m = PointNet_Plus(make_opt(256), precision="bf16")
    m.load_state_dict(load_states()["pointnet"], strict=False)
    m = m.to(dev).eval()
    f = m.folded()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    x1 = torch.empty((B, N1, 132), dtype=torch.float32, device=dev)
    ar = torch.arange(N, device=dev).expand(B, N)

    def step(timers=None):
        def timed(name, fn):
            if timers is None:
                return fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = fn(); b.record()
            timers.setdefault(name, []).append((a, b))
            return out
        order = timed("fps", lambda: ops.fps(pts, N1, start))
        def reorder():
            mask = torch.ones((B, N), dtype=torch.bool, device=dev)
            mask.scatter_(1, order.long(), False)
            rest = ar[mask].view(B, N - N1)
            return torch.gather(pts, 1, torch.cat([order.long(), rest], 1)[..., None].expand(-1, -1, 3)).contiguous()
        cloud = reorder()
        idx = timed("ball_query", lambda: ops.knn_ball(cloud, N1, K, 0.01))
        timed("point_mlp", lambda: m._sa(cloud, idx, "netR_1", f, x1, None))
        return x1

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    timers, evs = {}, []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(timers); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    per = {k: sorted(a.elapsed_time(b) for a, b in v)[len(v) // 2] for k, v in timers.items()}
    bytes_ = {"fps": 14336 * B, "ball_query": 143360 * B, "point_mlp": (12288 + 131072 + 512 * 132 * 4) * B}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    kern = {k: {"ms": round(per[k], 4), "hbm_gbs": round(bytes_[k] / (per[k] * 1e-3) / 1e9, 2),
                "hbm_frac": round(bytes_[k] / (per[k] * 1e-3) / 1e9 / peaks["hbm_gbs"], 5)} for k in per}
    kern["point_mlp"]["tflops"] = round(FLOP_SA1 * B / (per["point_mlp"] * 1e-3) / 1e12, 2)
    kern["point_mlp"]["tensor_frac"] = round(kern["point_mlp"]["tflops"] / peaks["bf16_tflops"], 4)
    rec = {
        "metric": "sa_microbench_clouds_per_sec", "value": B / (ms * 1e-3), "unit": "clouds/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg2-sa-microbench: %d clouds x 1024 pts, FPS 1024->512, ball query r2=0.01 k=64, "
                               "fused 3->64->64->128 point-MLP + max" % B, "l2": "256 MiB flush between steps"},
        "kernels": kern,
        "note": "FPS and the neighbour search keep the 12 KB cloud in shared memory/registers: they are SM "
                "latency/issue bound by construction, so their HBM fraction is tiny (SURVEY 8d); %d clouds "
                "occupy %d of 148 SMs in the one-CTA-per-cloud FPS kernel" % (B, min(B, 148)),
    }
    if emit:
        print(json.dumps(rec))
    return rec


def cfg5_workload(args):
    return ("cfg5-train-hotpath: %d frames/GPU x 2 hands, %dx%d pyramid, 1024-pt clouds; train-mode forward "
            "(batch-stat BatchNorm, per hand) + backward of pyramid gather/SFT0+SA1+SFT1+SA2+SFT2+global MLP+"
            "fusion SFT, gradient all-reduce (mean), Adam step" % (args.frames, args.res, args.res))


def cfg5_inputs(B, R, seed):
    from pdfnet_b200 import synth
    g = torch.Generator().manual_seed(seed)
    return dict(cloud=synth.clouds(2 * B, seed=seed).view(B, 2, 1024, 3),
                choose=synth.choose_indices(2 * B, R, seed=seed).view(B, 2, 1024),
                l0=synth.pyramid(B, R, seed=seed)[0], l1=synth.pyramid(B, R, seed=seed)[1],
                l2=synth.pyramid(B, R, seed=seed)[2], center=torch.randn((B, 2, 1024), generator=g),
                target=torch.randn((B, 2, 1024), generator=g))


def run_cfg5_reference(args):
    """CPU arm of cfg5: the oracle's train-mode restatement (torch-CPU autograd, fp32, all host
    threads) + Adam on a bounded sample of the same workload; rank 0 only."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle import pdf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    R, B = args.res, args.cpu_sample_frames
    opt = make_opt(R)
    inp = cfg5_inputs(B, R, 317)
    st = load_states()
    sd = {k: v.clone() for k, v in st["pointnet"].items()}
    sft = {k: v.clone().requires_grad_(True) for k, v in st["sft"].items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    params = [v for v in list(sd.values()) + list(sft.values()) if v.requires_grad]
    optim = torch.optim.Adam(params, lr=1e-4)
    emb = [inp["l0"], inp["l1"], inp["l2"]]

    def step():
        optim.zero_grad(set_to_none=True)
        l = O.pointnet_plus_train(sd, inp["cloud"][:, 0], emb, inp["choose"][:, 0], opt)
        r = O.pointnet_plus_train(sd, inp["cloud"][:, 1], emb, inp["choose"][:, 1], opt)
        fused = O.sft_layer(torch.cat((l, r), 1).transpose(1, 2), inp["center"], sft)
        loss = ((fused - inp["target"]) ** 2).mean()
        loss.backward()
        optim.step()
        return float(loss.detach())

    warm = min(args.warmup, 1)
    for _ in range(warm):
        step()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    val = B / (ms / 1e3)
    cores = os.cpu_count() or 1
    print(json.dumps({
        "impl": "reference", "metric": "train_frames_per_sec_hot_path", "value": val, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg5_workload(args), "sample_frames_per_step": B, "resolution": R},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d frames/step, oracle train-mode port (torch CPU autograd fp32, %d threads)"
                                   % (B, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_cfg5(args, emit=True, dist_ready=None):
    """BASELINE.json configs[4] restricted to the hot path: one training step = train-mode forward +
    backward on this repo's kernels, bucketed gradient all-reduce over NCCL overlapped with the backward
    pass, Adam update.  emit=False: called from the default workload, returns the record instead of printing."""
    from pdfnet_b200 import parallel
    torch.set_grad_enabled(True)
    rank, world, local = dist_ready if dist_ready is not None else parallel.init_distributed("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from pdfnet_b200 import HandFusion, _lib, training
    R, B = args.res, args.frames
    model = HandFusion(make_opt(R), precision=args.precision)
    st = load_states()
    sd = {"pointnet_plus." + k: v for k, v in st["pointnet"].items()}
    sd.update({"sft." + k: v for k, v in st["sft"].items()})
    model.load_state_dict(sd, strict=False)
    model = model.to(dev).train()
    params = [p for n, p in model.named_parameters() if not n.startswith("mano_head") and "netR_FC" not in n]
    optim = torch.optim.Adam(params, lr=1e-4, fused=True)
    host = cfg5_inputs(B, R, 317 + rank)
    pinned = {k: v.contiguous().pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    staging = {k: torch.empty_like(v) for k, v in resident.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    comm_bytes = [0]
    sync = training.BucketedAllReduce(params, world, bucket_bytes=args.bucket_mb << 20)

    def train_step(d):
        sync.zero_grad()                                       # gradients are views of the all-reduce buckets
        fused = model(d["cloud"], [d["l0"], d["l1"], d["l2"]], d["choose"], d["center"])
        loss = ((fused - d["target"]) ** 2).mean()
        loss.backward()                                        # hooks launch each bucket's all-reduce as it fills
        comm_bytes[0] = sync.finish()
        optim.step()
        return loss

    def e2e_step():
        for k, v in pinned.items():
            staging[k].copy_(v, non_blocking=True)
        return float(train_step(staging).detach())           # device -> host read of the loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        return parallel.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps, dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms_dev = timed_loop(lambda: train_step(resident), args.steps, args.warmup)
    launches = (_lib.launch_count() - l0) // (args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed_loop(e2e_step, args.steps, args.warmup)
    loss = float(train_step(resident).detach())
    peak_mem = torch.cuda.max_memory_allocated(dev)
    sync.remove()
    rec = None
    if rank == 0:
        n_clouds = 2 * B
        fwd = (FLOP_SA1 + FLOP_SA2 + FLOP_GLOBAL + FLOP_SFT1 + FLOP_SFT2) * n_clouds
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
        ach = 3 * fwd / (ms_dev * 1e-3) / 1e12                 # forward + data-gradient + weight-gradient GEMMs
        rec = {
            "metric": "train_frames_per_sec_hot_path", "value": world * B / (ms_dev * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split operands, fp32-accurate)", "data": "synthetic",
            "config": {"workload": cfg5_workload(args), "frames_per_gpu": B, "resolution": R,
                       "precision": args.precision + (" operands for the forward, dX and dW GEMMs (fp32 activations, fp32 accumulate)"
                                                      if args.precision == "bf16" else " (split-bf16 tensor-core GEMMs)"),
                       "parallelism": "dp%d" % world, "l2": "256 MiB flush write between timed iterations",
                       "optimizer": "Adam (torch fused), lr 1e-4; loss = MSE(fused, target)"},
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pinned.values()),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "whole step (forward + dX + dW GEMM flops over the step time)", "bound": "tensor",
                         "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": ach / peaks["bf16_tflops"], "traffic": None,
                         "pipe": "tcgen05 bf16; the step is dominated by HBM-bound staging / BatchNorm passes"},
            "allreduce_bytes_per_step": comm_bytes[0], "allreduce_buckets": len(sync.buckets),
            "allreduce": "bucketed (%d MiB), launched from gradient hooks, overlapped with the backward pass" % args.bucket_mb,
            "final_loss": loss, "peak_mem_gb": peak_mem / 2 ** 30, "cpu_baseline": None}
        if emit:
            print(json.dumps(rec))
    del model, optim, sync, resident, staging
    torch.cuda.empty_cache()
    if emit and world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return rec


def decoder_flops_per_frame():
    """Dense-layer + attention FLOPs of decoder.forward for one frame (two hands)."""
    cins, couts, verts = (512, 256, 128), (256, 128, 64), (63, 126, 252)
    fl = 2 * 1024 * 509                                                  # gf_layer
    for ci, co, V in zip(cins, couts, verts):
        g = 0
        for b in range(4):
            c0 = ci if b == 0 else co
            g += 2 * V * (c0 * 3 * co + co * 2 * co)                       # [W0;W1;shortcut], [W0';W1']
        sa = 2 * V * (3 * co * co + co * co + 2 * co * co) + 4 * V * V * co  # qkv, fc, ff + QK^T, PV
        ia = 2 * V * (3 * co * co + co * co + 2 * co * co) + 4 * V * V * co
        fl += g + sa + ia
    fl += 2 * 252 * 64 * 3 + 2 * 3 * 252 * 778
    return 2 * fl


def run_decoder(args):
    """the GCN decoder that consumes fuse_feat (the reference's live path after the
    fusion tail).  One step = decoder.forward for ``--frames`` frames (both hands), replayed as a CUDA graph."""
    from pdfnet_b200 import parallel
    torch.set_grad_enabled(False)
    rank, world, local = parallel.init_distributed("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from pdfnet_b200 import _lib, synth
    from pdfnet_b200.decoder import decoder
    from pdfnet_b200.graph import CapturedStep
    assets = dict(np.load(os.path.join(ROOT, "tests", "golden", "gcn_assets.npz")))
    B = args.frames
    prec = "fp32" if args.precision == "fp32" else "bf16x3"
    m = decoder(assets, precision=prec)
    state = synth.decoder_state(317, assets["upsample"])
    m.load_state_dict(state)
    m = m.to(dev).eval()
    fuse_host = torch.randn((B, 2, 1024), generator=torch.Generator().manual_seed(317 + rank)).pin_memory()
    fuse = fuse_host.to(dev)
    fl, fr = fuse[:, 0].contiguous(), fuse[:, 1].contiguous()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def fwd():
        result, params, _, _ = m(fl, fr, None)
        return result["verts3d"]["left"], result["verts3d"]["right"], params["root"]["left"], params["root"]["right"]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps, dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_eager = timed(fwd, args.steps, args.warmup)
    step = CapturedStep(fwd)
    ms_dev = timed(step.replay, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    out_host = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in step.outputs]

    def e2e():
        fuse.copy_(fuse_host, non_blocking=True)
        fl.copy_(fuse[:, 0]); fr.copy_(fuse[:, 1])
        outs = step.replay()
        for h, t in zip(out_host, outs):
            h.copy_(t, non_blocking=True)

    ms_e2e = timed(e2e, args.steps, args.warmup)
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
        flops = decoder_flops_per_frame() * B
        ach = flops / (ms_dev * 1e-3) / 1e12
        from oracle import pdf_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        ns = args.cpu_sample_frames
        sd = state
        with torch.no_grad():
            O.gcn_decoder_forward(sd, assets, fuse_host[:ns])
            t0 = time.perf_counter(); O.gcn_decoder_forward(sd, assets, fuse_host[:ns]); t_cpu = time.perf_counter() - t0
        cores = os.cpu_count() or 1
        print(json.dumps({
            "metric": "decoder_frames_per_sec", "value": world * B / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "eager_ms_per_step": ms_eager,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if prec == "fp32" else "bf16x3 (split operands, fp32-accurate)", "data": "synthetic",
            "config": {"workload": "f3-gcn-decoder: %d frames/GPU x 2 hands, fuse_feat [B,2,1024] -> 3 DualGraph levels "
                                   "(63/126/252 vertices, 4 GCN blocks + self/inter attention each) -> 778-vertex meshes"
                                   % B, "precision": prec, "launch": "one CUDA-graph replay per step",
                       "l2": "256 MiB flush write between timed iterations"},
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": fuse_host.numel() * 4,
                    "d2h_bytes_per_step": sum(h.numel() * 4 for h in out_host)},
            "gpu_launches": int(step.launches), "clocks": clocks,
            "roofline": {"kernel": "whole decoder step (%.2f GFLOP/frame of dense + attention math)"
                                   % (decoder_flops_per_frame() / 1e9), "bound": "tensor", "achieved": ach,
                         "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                         "traffic": None,
                         "pipe": "latency / launch-count bound: ~290 small kernels per step, none larger than 30 us"},
            "cpu_baseline": {"value": ns / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d frames, oracle port of decoder.forward (torch CPU fp32, %d threads)" % (ns, cores)}}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"],
                    help="cfg3 default bf16; cfg5 default fp32 = split-bf16 (3-pass, fp32-accurate) tensor-core GEMMs, "
                         "bf16 = plain bf16 operands for forward / dW")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (cfg3: 128, cfg5: 64)")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--cpu-sample-frames", type=int, default=16,
                    help="frames per CPU step of the reference arm / cpu_baseline leg (about 1.6 s of 16-thread CPU work each)")
    ap.add_argument("--pyramid", default=None, choices=["bf16-nhwc", "fp32-nchw", "fp32-nhwc"],
                    help="dtype / memory format of the RGB feature pyramid inputs: bf16-nhwc = what an autocast "
                         "channels-last neck emits (default in bf16 mode; BASELINE cfg3 is a bf16 forward), fp32-nchw = "
                         "what the reference neck emits (default in fp32 mode), fp32-nhwc = SURVEY 8f row f4")
    ap.add_argument("--masks", default="u8", choices=["u8", "f32"], help="hand mask dtype (f32 = the reference's)")
    ap.add_argument("--no-decoder", dest="with_decoder", action="store_false",
                    help="cfg3: leave the GCN decoder + joint regressor (SURVEY 8f row f3, a15) out of the step")
    ap.add_argument("--no-sub", action="store_true",
                    help="cfg3: skip the cfg2 (SA microbench) and cfg5 (training step) sub-records of the line")
    ap.add_argument("--sustained-seconds", type=float, default=2.0,
                    help="cfg3: also loop the step for this long and report value_sustained (0 = off)")
    ap.add_argument("--bucket-mb", type=int, default=4, help="cfg5: gradient all-reduce bucket size (MiB)")
    ap.add_argument("--pipeline", action="store_true",
                    help="cfg3: also time the step software-pipelined over consecutive batches (point branch of batch i beside "
                         "the GCN decoder of batch i-1, pdfnet_b200.graph.PipelinedStep) -> 'pipelined' sub-record; measured: no "
                         "gain (3.11 vs 3.15 ms, DESIGN 3.5), hence off by default")
    ap.add_argument("--no-graph", action="store_true", help="time the eager launch path instead of a CUDA-graph replay")
    ap.add_argument("--e2e-pyramid", default="zero-copy", choices=["zero-copy", "copy"],
                    help="e2e hand-off of the host-resident bf16 feature pyramid: zero-copy = the gather kernel reads the "
                         "selected pixels in place from page-locked host memory (default), copy = dense host->device copy "
                         "of the maps first; both are measured, the other one is reported as e2e_other_handoff")
    ap.add_argument("--e2e-zero-copy-levels", default="auto",
                    help="which pyramid maps the zero-copy hand-off reads in place (the rest is copied).  auto: l1,l2 on one "
                         "or two GPUs (l0's 6-byte pixels cost a PCIe read each: while the host feeds every link at full "
                         "speed the dense 50 MB copy of l0 is 2 - 8 %% faster than 262 k tiny reads), l0,l1,l2 from four GPUs "
                         "on, where the host's memory system is the limit (35 %% fewer bytes over PCIe per step)")
    ap.add_argument("--e2e-streams", type=int, default=2, help="compute streams the e2e chunks alternate between")
    ap.add_argument("--e2e-chunks", type=int, default=2, help="H2D/compute overlap chunks in the e2e measurement")
    ap.add_argument("--host-alloc", default="wc", choices=["wc", "pinned"],
                    help="e2e input staging buffers: write-combined page-locked memory (cudaHostAllocWriteCombined) or "
                         "torch's pin_memory()")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg2", "cfg5", "decoder"],
                    help="cfg3 = hot path at 128 frames/GPU (default, the driver's contract); cfg2 = SA microbench; "
                         "cfg5 = training step (use --frames 64)")
    ap.add_argument("--clouds", type=int, default=64, help="cfg2: number of clouds")
    args = ap.parse_args()
    if args.gpus > 1 and "RANK" not in os.environ and args.impl == "ours" and args.workload != "cfg2":
        # launched without torchrun: start one rank per GPU ourselves (same command line the driver uses)
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.frames is None:
        args.frames = 64 if args.workload == "cfg5" else 128
    if args.precision is None:
        args.precision = "bf16"                       # cfg5: BASELINE's "training step bf16"; --precision fp32 = the fp32-accurate mode
    if args.pyramid is None:
        args.pyramid = "bf16-nhwc" if args.precision == "bf16" else "fp32-nchw"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "cfg5":
        (run_cfg5_reference if args.impl == "reference" else run_cfg5)(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg2":
        run_cfg2(args)
    elif args.workload == "decoder":
        run_decoder(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
