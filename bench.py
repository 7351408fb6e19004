#!/usr/bin/env python3
"""bench.py — headline benchmark of the matching hot path (BASELINE.json metric).

Workload (configs[4] of BASELINE.json): AhoCorasickSet, 1,000,000 synthetic a-z keywords (len 3-12),
case-sensitive, all overlapping matches over a 64 GB corpus = 32 independent haystacks of 10^9 UTF-16 chars
(a Java String holds < 2^31 chars, so the corpus is necessarily many match() calls), generated ON THE DEVICE
from SplitMix64 seeds.  The corpus is sharded by haystack across the N ranks ("strong" scaling: total fixed);
a step = one match() of every haystack of the corpus.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
  python bench.py --impl reference ...                           # the reference's algorithm on the host cores

One JSON line on stdout (rank 0).  `value` = haystack GB/s with inputs resident in HBM (device-timed, CUDA
events on the launch stream, max over ranks); `e2e` = the same metric through the host-buffer C-ABI call
acgpu_match_utf16 (H2D copy, kernels, D2H of the match records inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import workloads as W  # noqa: E402

METRIC = "haystack_GB_per_s"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--haystacks", type=int, default=32, help="haystacks in the corpus (total over all ranks)")
    ap.add_argument("--chars", type=int, default=1_000_000_000, help="UTF-16 chars per haystack")
    ap.add_argument("--keywords", type=int, default=1_000_000)
    ap.add_argument("--e2e-chars", type=int, default=1_000_000_000, help="chars of the end-to-end haystack (default: a full haystack)")
    ap.add_argument("--no-replay", action="store_true", help="skip the e2e leg with the counting listener (C++ mirror)")
    ap.add_argument("--cpu-sample-chars", type=int, default=16_000_000, help="chars per host thread per CPU pass")
    ap.add_argument("--shard", default="haystack", choices=["haystack", "range"],
                    help="haystack: the corpus' haystacks are dealt to the ranks (default); range: EVERY haystack is cut by end-position "
                         "range across the ranks (plan_range_shards), each rank holds only its slice, the per-shard match counts are "
                         "all-gathered inside the timed region and the rank-ordered stream is checked against the single-GPU stream")
    ap.add_argument("--config", type=int, default=4, choices=[0, 1, 2, 3, 4, 5],
                    help="BASELINE.json configs[i] (default 4 = the headline: the metric is quoted on it); 0-3: the other configs, one "
                         "JSON line per matcher with the same keys; 5: the English-like 'real dictionary' workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args):
    return {
        "workload": "configs[4]: AhoCorasickSet, %d a-z keywords (len 3-12), case-sensitive, all overlapping "
                    "matches, %d haystacks x %d UTF-16 chars (%.1f GB corpus), sharded by haystack"
                    % (args.keywords, args.haystacks, args.chars, args.haystacks * args.chars * 2 / 1e9),
        "keywords": args.keywords, "haystacks": args.haystacks, "chars_per_haystack": args.chars,
        "l2": "inputs larger than L2 (each haystack is %.1f GB; no flush needed)" % (args.chars * 2 / 1e9),
        "dict_seed": 1005, "haystack_seed": "2005+i",
    }


# ----------------------------------------------------------------------------- CPU arm (oracle on host cores)

def cpu_pass(matcher, slices, results):
    threads = []
    for i, sl in enumerate(slices):
        def run(i=i, sl=sl):
            results[i] = matcher.count(sl)
        t = threading.Thread(target=run)
        threads.append(t)
        t.start()
    for t in threads:
        t.join()


def cpu_measure(args, keywords, spec, passes: int, warm: int, gpu_matcher=None):
    """All host threads, one independent match() per thread over disjoint slices of haystack 0 (the only
    parallelism the reference API allows).  ctypes releases the GIL inside the oracle call."""
    from oracle import oracle as ora
    cores = os.cpu_count() or 1
    m = ora.Matcher("ahocorasick", keywords, case_sensitive=True)
    n = args.cpu_sample_chars
    slices = [W.make_haystack(spec, n, start=i * ((n + 63) // 64 * 64)) for i in range(cores)]
    res = [0] * cores
    for _ in range(warm):
        cpu_pass(m, slices, res)
    t0 = time.perf_counter()
    for _ in range(passes):
        cpu_pass(m, slices, res)
    dt = (time.perf_counter() - t0) / passes
    gbps = cores * n * 2 / dt / 1e9
    parity = None
    if gpu_matcher is not None:
        # post-timing parity spot check (the oracle as the CHECKER): the CUDA path through the host-buffer C-ABI call on
        # the very slices the CPU arm just scanned - match counts of every slice, the full ordered record stream of slice 0
        got = [len(gpu_matcher.match_records(sl)) for sl in slices]
        want0 = m.match(slices[0], cap=n)
        rec0 = gpu_matcher.match_records(slices[0])
        ok = got == [int(r) for r in res] and len(rec0) == len(want0) and \
            bool(np.array_equal(rec0.start, want0["start"])) and bool(np.array_equal(rec0.end, want0["end"]))
        parity = {"checked": True, "ok": ok, "slices": cores, "chars_per_slice": n, "records_compared": int(len(want0)),
                  "counts_compared": int(sum(res))}
        if not ok:
            raise SystemExit("bench.py: parity spot check FAILED (CUDA path != oracle): counts %r vs %r" % (got, res))
    return dict(value=gbps, unit=UNIT, cores=cores, kind="port", parity=parity,
                sample="%d threads x %d chars of haystack 0 per pass (AhoCorasickSet, same dictionary); "
                       "literal C restatement of the reference (no JVM in this image), gcc -O2" % (cores, n),
                matches_per_s=sum(res) / dt, ms_per_pass=dt * 1e3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = W.config(4, scale=args.keywords / 1_000_000)
    kws = cfg["keywords"]
    spec = W.HaystackSpec("lower", 2005, kws)
    r = cpu_measure(args, kws, spec, passes=args.steps, warm=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_pass"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": workload_config(args), "matches_per_s": r["matches_per_s"],
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- e2e with listener replay (C++ mirror)

def replay_measure(kws, host, ne, device, steps):
    """AhoCorasickSet.match(String, SetMatchListener) through include/acgpu.hpp with a counting listener (tools/e2e_replay.cpp)."""
    from ahocorasick_b200 import _lib
    src = os.path.join(ROOT, "tools", "e2e_replay.cpp")
    so = os.path.join(ROOT, "tools", "libe2e_replay.so")
    libdir = os.path.dirname(_lib.LIB_PATH)
    deps = [src, os.path.join(ROOT, "include", "acgpu.hpp"), os.path.join(ROOT, "include", "acgpu.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", src, "-I" + os.path.join(ROOT, "include"),
                               "-L" + libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir, "-o", so])
    shim = C.CDLL(so)
    shim.e2e_replay_count.restype = C.c_double
    shim.e2e_replay_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_int64)]
    chars, offsets = W.keywords_to_arrays(kws)
    chars = np.ascontiguousarray(chars)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    matches = C.c_int64(0)
    secs = shim.e2e_replay_count(chars.ctypes.data, offsets.ctypes.data, len(kws), host.data_ptr(), ne, device, 1, steps,
                                 C.byref(matches))
    if secs < 0:
        raise RuntimeError("e2e_replay_count failed")
    return {"value": ne * 2 / secs / 1e9, "unit": UNIT, "ms_per_call": secs * 1e3, "listener_calls_per_call": matches.value,
            "steps": steps, "note": "acgpu::AhoCorasickSet::match(String, counting listener) - include/acgpu.hpp; one host thread "
                                    "replays every match"}


# ----------------------------------------------------------------------------- clocks sampler

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    import ahocorasick_b200 as ac
    from ahocorasick_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the matching path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lib = _lib.lib()
    cfg = W.config(4, scale=args.keywords / 1_000_000)
    kws = cfg["keywords"]
    t0 = time.perf_counter()
    matcher = ac.AhoCorasickSet(kws, True, device=local_rank)
    build_s = time.perf_counter() - t0
    info = matcher.info()

    # corpus shard of this rank: haystacks rank, rank+world, ... generated on the device
    mine = list(range(rank, args.haystacks, world))
    n = args.chars
    hays = []
    for i in mine:
        spec = W.HaystackSpec("lower", 2005 + i, kws)
        hays.append(W.make_haystack_torch(spec, n, device=dev))
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    # size the record buffer from a counting run (cap = 0 writes nothing, the kernel still counts)
    counts = []
    for h in hays:
        tot = C.c_int64(0)
        _lib.check(lib.acgpu_match_device(matcher.handle, h.data_ptr(), n, 0, n, None, None, 0, C.byref(tot), sp))
        counts.append(tot.value)
    cap = max(counts) if counts else 0
    d_pos = torch.empty((max(cap, 1), 2), dtype=torch.int32, device=dev)
    d_tot = torch.zeros(max(len(hays), 1), dtype=torch.int64, device=dev)
    launches = lib.acgpu_launches_per_match(matcher.handle)

    def step():
        for j, h in enumerate(hays):
            _lib.check(lib.acgpu_match_device_async(matcher.handle, h.data_ptr(), n, 0, n, d_pos.data_ptr(), None, cap,
                                                    d_tot[j:].data_ptr(), sp))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    got = d_tot[:len(hays)].tolist()
    assert got == counts, "match counts changed between runs: %r vs %r" % (got, counts)

    # per-shard match counts: the one exchange the path has — all_gather over NCCL
    my_matches = torch.tensor([sum(counts)], dtype=torch.int64, device=dev)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(my_matches) for _ in range(world)]
        dist.all_gather(gathered, my_matches)
        total_matches = int(sum(int(g.item()) for g in gathered))
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    else:
        total_matches = int(my_matches.item())
    ms_max = float(t_ms.item())
    ms_per_step = ms_max / args.steps
    total_chars = args.haystacks * n
    value = total_chars * 2 / (ms_per_step * 1e-3) / 1e9
    matches_per_s = total_matches / (ms_per_step * 1e-3)

    # roofline of one match = k_tier_mask + k_row_scan + k_tier_emit back to back (the scan is 0.4 % of it):
    # algorithmic bytes per match (2 B per char read once + 8 B per record written once) / average duration of the
    # three launches, timed by the CUDA events above on the launch stream
    n_launch = max(len(hays), 1)
    alg_bytes = 2 * n + 8 * (sum(counts) / n_launch)
    launch_ms = ms / args.steps / n_launch
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    # DRAM bytes of the same three launches from the committed ncu --set full captures (profiles/traffic.json), scaled
    # to this haystack length; null when the captures are for another workload
    traffic, traffic_note = None, "no capture"
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import make_traffic
        if t.get("keywords") != args.keywords:
            traffic_note = "capture is for another dictionary"
        elif t.get("kernel_sources_sha") != make_traffic.kernel_sources_sha():
            traffic_note = "stale: the kernels changed since the ncu capture (re-run tools/make_traffic.py)"
        else:
            traffic = t["dram_bytes_per_char"] * n
            traffic_note = "ncu --set full capture of these kernel sources (%s)" % t.get("source", "")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note, "kernel": "k_tier_mask + k_row_scan + k_tier_emit (one match)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": launch_ms,
                "haystack_only_frac": (2 * n / (launch_ms * 1e-3) / 1e9) / peak,
                "frac_of_nominal_8000": achieved / 8000.0}  # also against the nominal HBM3e figure

    # end-to-end through the host-buffer C-ABI call a user of the mirrors makes (acgpu_match_utf16_compact): H2D of the
    # haystack from pinned memory, kernels, D2H of the match stream, all inside the timed region; one FULL haystack of the
    # config per call, --steps calls after 2 warm-up calls
    e2e = None
    if not args.no_e2e:
        ne = min(n, args.e2e_chars, 0x7FFFFFFF)
        host = torch.empty(ne, dtype=torch.int16, pin_memory=True)
        host.copy_(hays[0][:ne] if hays else W.make_haystack_torch(W.HaystackSpec("lower", 2005, kws), ne, device=dev))
        torch.cuda.synchronize()
        res = _lib.Matches()
        times, n_rec, kind = [], 0, 0
        e2e_steps = max(1, args.steps)
        for it in range(2 + e2e_steps):
            barrier()
            t0 = time.perf_counter()
            _lib.check(lib.acgpu_match_utf16_compact(matcher.handle, host.data_ptr(), ne, C.byref(res)))
            dt = time.perf_counter() - t0
            n_rec, kind = int(res.n), int(res.kind)
            lib.acgpu_free_matches(C.byref(res))
            if it >= 2:
                times.append(dt)
        t_e = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        d2h = ne * 2 if kind == _lib.MATCHES_MASKS else n_rec * 8
        e2e = {"value": world * ne * 2 / float(t_e.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": ne * 2,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_call": float(t_e.item()) * 1e3,
               "matches_per_call": n_rec, "wire_format": "hit masks, 2 B/char" if kind == _lib.MATCHES_MASKS else "records, 8 B/match",
               "note": "acgpu_match_utf16_compact on one %d-char haystack per rank from pinned host memory, match stream "
                       "copied back to pinned host memory; mean of %d calls after 2 warm-ups (max over ranks)" % (ne, e2e_steps)}
        # the same call through the C++ mirror of the reference API with a counting SetMatchListener: + one listener call
        # per match on one host thread (what the CPU arm pays inside its loop)
        if rank == 0 and world == 1 and not args.no_replay:
            try:
                e2e["with_replay"] = replay_measure(kws, host, ne, local_rank, max(1, min(2, args.steps)))
            except Exception as exc:  # the shim needs g++; the headline does not depend on it
                e2e["with_replay"] = {"error": str(exc)[:200]}
        del host

    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_measure(args, kws, W.HaystackSpec("lower", 2005, kws), passes=2, warm=1, gpu_matcher=matcher)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        parity = r["parity"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic", "config": workload_config(args),
            "matches_per_s": matches_per_s, "matches_per_step": total_matches,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * len(hays) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
            "parity_checked": bool(parity and parity["ok"]), "parity": parity,
            "dictionary": dict(info, build_seconds=build_s),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_range(args):
    """--shard range (SURVEY 8e / north_star: "large corpora shard by byte range across the GPUs of one box with boundary
    overlap; NCCL only all-gathers per-shard match counts and offsets").  Every haystack of the corpus is ONE match() whose
    end positions are cut into `world` ranges; rank r generates and holds only chars [read_from, emit_to) of each haystack."""
    import torch
    import torch.distributed as dist
    import ahocorasick_b200 as ac
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.sharding import plan_range_shards

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    cfg = W.config(4, scale=args.keywords / 1_000_000)
    kws = cfg["keywords"]
    matcher = ac.AhoCorasickSet(kws, True, device=local_rank)
    max_len = matcher.info()["max_len"]
    n = args.chars
    shard = plan_range_shards(n, world, max_len, align=64)[rank]
    s0 = shard.read_from // W.BLOCK * W.BLOCK        # the generator starts at multiples of its block; extra chars are context only
    n_loc = shard.emit_to - s0
    hays = [W.make_haystack_torch(W.HaystackSpec("lower", 2005 + i, kws), n_loc, start=s0, device=dev) for i in range(args.haystacks)]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    counts = []
    for h in hays:
        tot = C.c_int64(0)
        _lib.check(lib.acgpu_match_device(matcher.handle, h.data_ptr(), n_loc, shard.emit_from - s0, shard.emit_to - s0, None, None, 0,
                                          C.byref(tot), sp))
        counts.append(tot.value)
    cap = max(counts + [1])
    d_pos = torch.empty((cap, 2), dtype=torch.int32, device=dev)
    d_tot = torch.zeros(len(hays), dtype=torch.int64, device=dev)
    gathered = torch.zeros((len(hays), world), dtype=torch.int64, device=dev)

    def step():
        for j, h in enumerate(hays):
            _lib.check(lib.acgpu_match_device_async(matcher.handle, h.data_ptr(), n_loc, shard.emit_from - s0, shard.emit_to - s0,
                                                    d_pos.data_ptr(), None, cap, d_tot[j:].data_ptr(), sp))
            if world > 1:  # every rank learns every shard's count (its global record offset = the sum of the lower ranks')
                dist.all_gather_into_tensor(gathered[j], d_tot[j:j + 1])
            else:
                gathered[j] = d_tot[j:j + 1]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    per_shard = gathered.cpu().tolist()
    assert [row[rank] for row in per_shard] == counts, "match counts changed between runs"
    total_matches = int(sum(sum(row) for row in per_shard))

    # ---- the rank-ordered concatenation is the single-GPU stream: a position-weighted checksum of the LAST haystack's
    #      records (positions shifted to the whole haystack, offsets from the gathered counts) against rank 0 scanning
    #      that whole haystack alone
    def checksum(pos, first_index, shift):
        if pos.shape[0] == 0:
            return torch.zeros(3, dtype=torch.int64, device=dev)
        idx = torch.arange(first_index, first_index + pos.shape[0], dtype=torch.int64, device=dev)
        st, en = pos[:, 0].long() + shift, pos[:, 1].long() + shift
        mix = (st * 1_000_003 + en * 7_919 + 1) * ((idx % 1_000_033) + 1)   # wraps mod 2^64: order-sensitive
        return torch.stack([torch.tensor(pos.shape[0], dtype=torch.int64, device=dev), mix.sum(), (st ^ (en << 20)).sum()])

    last = len(hays) - 1
    offset = int(sum(per_shard[last][:rank]))
    mine = checksum(d_pos[:counts[last]], offset, s0)
    if world > 1:
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    check = None
    if rank == 0:
        del hays[:-1]
        whole = W.make_haystack_torch(W.HaystackSpec("lower", 2005 + last, kws), n, device=dev)
        tot = C.c_int64(0)
        _lib.check(lib.acgpu_match_device(matcher.handle, whole.data_ptr(), n, 0, n, None, None, 0, C.byref(tot), sp))
        d_all = torch.empty((max(tot.value, 1), 2), dtype=torch.int32, device=dev)
        _lib.check(lib.acgpu_match_device(matcher.handle, whole.data_ptr(), n, 0, n, d_all.data_ptr(), None, tot.value, C.byref(tot), sp))
        ref = checksum(d_all[:tot.value], 0, 0)
        check = bool(torch.equal(ref, mine))
        if not check:
            raise SystemExit("bench.py --shard range: the sharded stream differs from the single-GPU stream: %r vs %r" % (mine.tolist(), ref.tolist()))
    if rank == 0:
        total_chars = args.haystacks * n
        line = {
            "metric": METRIC, "value": total_chars * 2 / (ms_per_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": dict(workload_config(args), sharding="every haystack cut by END-position range across %d ranks (boundaries aligned "
                           "to 64 chars, %d chars of left context per shard); one NCCL all-gather of the shard counts per haystack INSIDE "
                           "the timed region" % (world, max_len - 1)),
            "matches_per_s": total_matches / (ms_per_step * 1e-3), "matches_per_step": total_matches, "clocks": clocks,
            "gpu_launches": lib.acgpu_launches_per_match(matcher.handle) * args.haystacks * args.steps,
            "collectives_per_step": args.haystacks if world > 1 else 0,
            "range_shard_stream_equals_single_gpu_stream": check, "shard_counts_last_haystack": per_shard[last],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


CONFIG_NAMES = {
    0: "configs[0]: AhoCorasickSet, 1 000 ASCII keywords (len 3-12), case-sensitive, all overlapping matches, 8e6 chars",
    1: "configs[1]: AhoCorasickMap(dict, dict, false), 100k keywords case-insensitive, 5e8 chars of mixed-case text with Latin-1 / Greek / Cyrillic letters",
    2: "configs[2]: LongestMatchMap and ShortestMatchSet, 100k nested keywords, leftmost non-overlapping selection, 2e9 chars",
    3: "configs[3]: WholeWordMatchSet / Map with custom word chars ['_','='], 50k keywords, 2e9 chars of mixed-punctuation text (Map also via Readable)",
    5: "real dictionary (not a BASELINE config): AhoCorasickSet / Map, 236 000 English-like words (1-24 chars, mixed case, apostrophes), case-sensitive, 1e9 chars",
}


def run_config(args):
    """configs[0..3] (and 5) on one GPU with the same JSON keys as the headline line; one line per matcher."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs
    from oracle import oracle as ora
    cfg = W.config(args.config)
    cores = os.cpu_count() or 1
    for d in bench_configs.measure([args.config], steps=args.steps, warmup=args.warmup):
        cpu = None
        if not args.no_cpu_baseline:
            # the oracle on one host thread over a bounded sample of the same workload (the families other than AhoCorasick
            # have no count-only entry: the sample is small)
            fam = {"AhoCorasick": "ahocorasick", "Longest": "longest", "Shortest": "shortest", "WholeWordLongest": "wholewordlongest",
                   "WholeWord": "wholeword"}[next(k for k in ("WholeWordLongest", "WholeWord", "AhoCorasick", "Longest", "Shortest") if d["matcher"].startswith(k))]
            n_s = 4_000_000
            sample = W.make_haystack(cfg["spec"], n_s)
            kw = dict(case_sensitive=cfg["cs"])
            if "word_chars" in cfg and fam.startswith("wholeword"):
                kw["word_chars_table"] = ora.word_chars(2, *cfg["word_chars"])
            try:
                om = ora.Matcher(fam, cfg["keywords"], n_values=len(cfg["keywords"]) if "Map" in d["matcher"] else -1, **kw)
                t0 = time.perf_counter()
                om.match(sample, cap=2 * n_s)
                dt = time.perf_counter() - t0
                cpu = {"value": n_s * 2 / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                       "sample": "1 thread x %d chars (literal C restatement of the reference loop; no JVM in this image)" % n_s}
            except ora.OracleError:
                cpu = None
        line = {"metric": METRIC, "value": d["haystack_GB_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": d["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
                "config": {"workload": CONFIG_NAMES[args.config], "matcher": d["matcher"], "chars": d["chars"], "keywords": d["keywords"],
                           "l2": "inputs larger than L2"},
                "matches_per_s": d["matches_per_s"], "matches_per_step": d["matches"],
                "e2e": {"value": d["e2e_GB_per_s"], "unit": UNIT, "h2d_bytes_per_step": d["e2e_chars"] * 2, "d2h_bytes_per_step": d["e2e_d2h_bytes"],
                        "note": "acgpu_match_utf16 on a %d-char haystack from pinned host memory, best of 2 after 1 warm-up" % d["e2e_chars"],
                        "readable_stream_GB_per_s": d["readable_stream_GB_per_s"]},
                "gpu_launches": d["launches_per_match"] * args.steps, "roofline": dict(d["roofline"], unit="GB/s", traffic=None,
                                                                                        kernel="all launches of one match"),
                "cpu_baseline": cpu, "dictionary": d["info"]}
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config != 4:
        run_config(args)
    elif args.shard == "range":
        run_range(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
