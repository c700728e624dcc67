#!/usr/bin/env python
"""bench.py -- track-samples/s of the per-track analog scan on synthetic 9-track NRZI TBIN.

Metric (BASELINE.json): track-samples/s = rows x tracks scanned / time, with achieved HBM GB/s
against the measured roofline.  Workload at N=1: BASELINE.json configs[1] -- "synthetic 9-track
781 kHz TBIN, 10 Gsample, NRZI, 1 GPU, 1 parmset" = 10e9 track-samples = 1 111 111 111 rows x 9 x
int16 = 20.0 GB.  N>1: one process per GPU, each scans its own tape of that size
(time-sharded segments of an N-times longer reel, cut in inter-block gaps): weak scaling, no
data-path collective; NCCL carries only the barrier, the max-over-ranks time and the result gather
(per-rank event counts).

A step = one pass of the hot path over the whole tape:
  value : TBIN rows already resident in HBM -> ingest (de-interleave + quiet map) -> unit table ->
          scan kernel -> events in HBM                         [rt_attach_device + rt_bulk_scan]
  e2e   : the same through the C-ABI with HOST buffers: pinned host rows -> H2D -> ... -> events and
          proof data copied back to pinned host memory         [rt_bulk_scan_host: the three stages overlapped]
`--impl reference` times the reference's own CPU implementation (oracle/_ref/readtape_ref, the
unmodified readtape 3.18 binary built by oracle/Makefile) on all host cores, each step a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from readtape_b200 import abi, parmsets, synth, tbin, verify  # noqa: E402

FULL_ROWS = 1_111_111_111          # 10e9 track-samples / 9 tracks
METRIC = "track-samples/s, 9-track 781 kHz NRZI TBIN scan"
UNIT = "track-samples/s"


def workload_name(rows):
    return (f"synthetic 9-track 781.25 kHz 800 BPI NRZI TBIN, {rows * 9 / 1e9:.2f}e9 track-samples "
            f"({rows} rows, {rows * 18 / 1e9:.2f} GB int16), 1 parmset (NRZI #0), per GPU")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_memory_near_gpu(index):
    """Prefer host memory on the NUMA node the GPU hangs off for everything this process allocates from now on (the pinned tape and
    event buffers of the e2e run): with all ranks' buffers on one node, the GPUs of the other socket copy across the inter-socket
    link.  set_mempolicy(MPOL_PREFERRED) through libc; returns what was done, for the bench line.  Never fatal."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"gpu_pci": bus, "numa_node": node, "policy": "none (no NUMA information)"}
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238                   # x86-64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        return {"gpu_pci": bus, "numa_node": node, "policy": "MPOL_PREFERRED" if rc == 0 else f"set_mempolicy failed (errno {ctypes.get_errno()})"}
    except Exception as exc:
        return {"policy": f"unavailable ({exc!r})"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.lines, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- the reference CPU arm --------------------------------------------------------------------------
def reference_run(tile, ntiles_per_proc, nproc, steps, warmup, workdir):
    """nproc independent readtape_ref processes, each decoding its own TBIN of ntiles_per_proc super-tiles."""
    exe = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
    kind = "reference"
    if not os.path.exists(exe):
        return None
    hdr = synth.nrzi_header()
    path = os.path.join(workdir, "sample.tbin")
    rows = np.concatenate([tile] * ntiles_per_proc) if ntiles_per_proc > 1 else tile
    tbin.write_tbin(path, hdr, rows)
    nrows = rows.shape[0]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([exe, "-q", "-nm", "-nrzi", "-bpi=800", "-ips=50", "-tap", "-nolog", "-nolabels",
                                   f"-outf={workdir}/out{p}", path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                 for p in range(nproc)]
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError(f"readtape_ref exited {rcs}")
        if it >= warmup:
            times.append(dt)
    tsamp = nrows * 9 * nproc
    tap = os.path.join(workdir, "out0.tap")
    return {"value": tsamp * len(times) / sum(times), "ms_per_step": 1e3 * sum(times) / len(times), "kind": kind,
            "cores": nproc, "rows_per_proc": nrows, "tap_bytes": os.path.getsize(tap) if os.path.exists(tap) else None,
            "sample": f"{nproc} processes x {ntiles_per_proc} super-tiles ({nrows} rows x 9 tracks each) of the same synthetic "
                      f"tape, unmodified readtape 3.18 (gcc -O2), whole program incl. file read and .tap write"}


# ---- the product, end to end: TBIN file -> .tap through readtape with the B200 scan -------------------------------------------
def product_run(tile, workdir, gpus, nproc, reps=889, ref_tiles=56, runs=2):
    """readtape_b200 (the reference's own host code with readblock() replaced, readtape_b200/host) on a reel of `reps` super-tiles
    in the page cache, split between `nproc` worker processes (RT_WORKERS; one scanning parent on one GPU), wall clock from
    exec to exit; beside it the unmodified reference doing the same work with all host cores (`nproc` processes x `ref_tiles`
    super-tiles each, >= the same number of rows); the product's .tap must be the reference's."""
    exe = os.path.join(ROOT, "readtape_b200", "bin", "readtape_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "readtape_ref")
    if not (os.path.exists(exe) and os.path.exists(ref)):
        return {"unavailable": "readtape_b200/bin/readtape_b200 or oracle/_ref/readtape_ref not built"}
    hdr = tbin.build_header(synth.nrzi_header())
    end = np.array([tbin.END_MARK], dtype="<i2").tobytes()
    reel = os.path.join(workdir, "reel.tbin")
    with open(reel, "wb") as fh:
        fh.write(hdr)
        for _ in range(reps):
            tile.tofile(fh)
        fh.write(end)
    opts = ["-q", "-nm", "-nrzi", "-bpi=800", "-ips=50", "-tap", "-nolog", "-nolabels"]
    env = dict(os.environ, RT_WORKERS=str(nproc), RT_STATS="2")
    times, out = [], ""
    for _ in range(runs):
        t0 = time.perf_counter()
        r = subprocess.run([exe] + opts + [f"-outf={workdir}/prod", reel], capture_output=True, text=True, env=env)
        times.append(time.perf_counter() - t0)
        out = r.stdout
        if r.returncode != 0:
            return {"error": f"readtape_b200 exited {r.returncode}: {r.stdout[-400:]} {r.stderr[-400:]}"}
    phases = {"replay_s": 0.0, "events": 0, "hits": 0, "misses": 0, "restarts": 0}
    import re
    for m in re.finditer(r"(\d+) events, (\d+) speculative hits, (\d+) misses, (\d+) restarts", out):
        for key, v in zip(("events", "hits", "misses", "restarts"), m.groups()):
            phases[key] += int(v)
    for m in re.finditer(r"([\d.]+) s replaying events", out):
        phases["replay_s"] = max(phases["replay_s"], float(m.group(1)))             # the slowest worker
    for pat, key in ((r"rt_open ([\d.]+) s", "cuda_init_s"), (r"upload of [\d.]+ GB ([\d.]+) s", "file_to_gpu_s"),
                     (r"whole-tape scan ([\d.]+) s", "scan_s"), (r"results to the host ([\d.]+) s", "results_to_host_s")):
        m = re.search(pat, out)
        if m:
            phases[key] = float(m.group(1))
    # the reference, same work, all cores
    sample = os.path.join(workdir, "ref_sample.tbin")
    with open(sample, "wb") as fh:
        fh.write(hdr)
        for _ in range(ref_tiles):
            tile.tofile(fh)
        fh.write(end)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([ref] + opts + [f"-outf={workdir}/ref{p}", sample], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for p in range(nproc)]
    rcs = [p.wait() for p in procs]
    t_ref = time.perf_counter() - t0
    if any(rcs):
        return {"error": f"readtape_ref exited {rcs}"}
    # .tap identity: every block decode is independent and the reel is periodic, so the reel's .tap is the per-tile record
    # sequence repeated; the per-tile sequence comes from the reference's own output
    body = open(f"{workdir}/ref0.tap", "rb").read()
    assert body[-4:] == b"\xff" * 4
    body = body[:-4]
    per_tile = body[: len(body) // ref_tiles]
    identical = False
    if len(body) % ref_tiles == 0 and per_tile * ref_tiles == body:
        got = open(f"{workdir}/prod.tap", "rb").read()
        identical = len(got) == len(per_tile) * reps + 4 and got[-4:] == b"\xff" * 4 and all(
            got[i * len(per_tile):(i + 1) * len(per_tile)] == per_tile for i in range(reps))
    rows = reps * tile.shape[0]
    best = min(times)
    res = {"workload": f"TBIN file in the page cache ({reps} super-tiles, {rows} rows, {rows * 18 / 1e9:.2f} GB) -> .tap, whole program "
                       f"(readtape's own host code + the B200 scan), {nproc} worker processes on {gpus} GPU(s)",
           "seconds": best, "seconds_all_runs": times, "value": rows * 9 / best, "unit": UNIT, "tap_bytes": os.path.getsize(f"{workdir}/prod.tap"),
           "tap_identical_to_reference": bool(identical), "phases": phases,
           "reference": {"seconds": t_ref, "value": nproc * ref_tiles * tile.shape[0] * 9 / t_ref, "processes": nproc,
                         "super_tiles_per_process": ref_tiles, "what": "unmodified readtape 3.18 (gcc -O2), the same command line"},
           }
    res["ratio_vs_reference_all_cores"] = res["value"] / res["reference"]["value"]
    return res


def bench_capture(args, K, W):
    """BASELINE configs 3 and 5 on the reference's own captures (1 GPU):
      pe : examples/9trk_PE/LJS009_part1_39blks, all 8 built-in PE parameter sets scanned by ONE rt_bulk_scan call (fan-out);
           value = rows x 9 tracks x 8 passes / time, tape resident in HBM; plus the whole program (readtape_b200 with RT_FANOUT=1)
           beside the unmodified reference, .tap identical to the reference-held golden;
      ww : examples/6trk_Whirlwind/132_pt1 (full reel), the exact stateful scan (rt_scan_*: the detector state persists across
           blocks) driven through the reference's own reset sequence; plus the whole program beside the reference."""
    import hashlib
    import torch
    from oracle import captures
    from readtape_b200 import evlog
    torch.cuda.set_device(0)
    lib = abi.load_product()
    name = {"pe": "LJS009_part1_39blks", "ww": "132_pt1"}[args.workload]
    cap = captures.full_path(name)
    if cap is None:
        print(json.dumps({"metric": METRIC, "unavailable": f"capture {name} not staged"})); return
    doc, segs = evlog.load_fixture(os.path.join(ROOT, "tests", "golden", name + ".segments.json"))
    heads = doc["heads"]
    _, rows = tbin.read_tbin(cap, nheads=heads["nheads"])
    rows = np.ascontiguousarray(rows)
    desc = evlog.desc_from_heads(heads)
    tape = lib.open(desc)
    tape.upload(rows)
    nrows = tape.nrows
    full = json.load(open(os.path.join(ROOT, "tests", "golden", "full_outputs.json")))[name]
    sampler = ClockSampler(0)
    if args.workload == "pe":
        base = [s for s in segs if s.reset_kind == abi.RT_RESET_FULL and not (s.flags & abi.RT_F_DENSITY_DETECT) and s.parmset == 0][0]
        cfgs = [abi.make_cfg(base.mode, parmsets.PE[p], base.bpi, base.ips, flags=base.flags, skew=base.skew) for p in range(8)]
        for _ in range(W):
            b = tape.bulk_scan(cfgs); b.free()
        sampler.start(); time.sleep(0.25)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(K):
            b = tape.bulk_scan(cfgs); st = b.stats(); b.free()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
        passes, launches, events = 8, int(st.launches), int(st.events)
        what = "rt_bulk_scan, 8 PE parameter sets fanned out in one call (tape resident in HBM)"
    else:
        # the whole reset sequence of the reference's run (deskew pre-pass + main pass), cut to this capture's rows
        def run_all():
            n = 0
            for seg, got in evlog.replay(tape, segs):
                n += len(got)
            return n
        for _ in range(max(1, W // 3)):
            run_all()
        sampler.start(); time.sleep(0.25)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(K):
            events = run_all()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
        scanned = sum((s.end_row if s.end_row >= 0 else nrows) - s.row for s in segs)
        passes, launches = scanned / float(nrows), 0
        what = "rt_scan_* (exact stateful scan with skip-ahead) through the reference's reset sequence: deskew pre-pass + main pass"
    clocks = sampler.stop()
    tsamp = nrows * desc.ntrks * passes
    # whole program beside the reference
    work = tempfile.mkdtemp(prefix="rtcap_")
    prog = {}
    try:
        opts = [o for o in full["options"].split() if o not in ("-v", "-v3")]
        for tag, exe, env in (("readtape_b200", os.path.join(ROOT, "readtape_b200", "bin", "readtape_b200"), {"RT_FANOUT": "1", "RT_STATS": "1"}),
                              ("reference", os.path.join(ROOT, "oracle", "_ref", "readtape_ref"), {})):
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([exe] + opts + [f"-outf={work}/{tag}", cap], capture_output=True, text=True, env=dict(os.environ, **env), cwd=work)
                d = time.perf_counter() - t0
                best = d if best is None else min(best, d)
            prog[tag] = {"seconds": best, "rc": r.returncode}
            if tag == "readtape_b200":
                import re
                m = re.search(r"([\d.]+) s opening \+ upload, ([\d.]+) s in the scan library, ([\d.]+) s replaying", r.stdout)
                if m:
                    prog[tag].update(open_upload_s=float(m.group(1)), scan_library_s=float(m.group(2)), replay_s=float(m.group(3)))
        ok = True
        for fname, want in full["outputs"].items():
            for tag in ("readtape_b200", "reference"):
                pth = os.path.join(work, fname.replace(name, tag, 1))
                ok = ok and os.path.exists(pth) and hashlib.sha256(open(pth, "rb").read()).hexdigest() == want["sha256"]
        prog["outputs_identical_to_reference_golden"] = bool(ok)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    peak, peak_src = measured_peak()
    line = {"metric": METRIC.replace("9-track 781 kHz NRZI TBIN", f"{name} ({'9-track PE, 8 parameter sets' if args.workload == 'pe' else '6-track Whirlwind'})"),
            "value": tsamp / dt, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "reference capture " + name,
            "config": {"workload": f"{name}.tbin, {nrows} rows x {desc.ntrks} tracks, {passes:.2f} scan passes: {what}", "events": int(events)},
            "roofline": {"bound": "hbm", "kernel": "k_units_sparse / k_peak_masks x 8" if args.workload == "pe" else "k_ctx_scan", "achieved": 2.0 * tsamp / dt / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": 2.0 * tsamp / dt / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "note": "a 40-80 MB capture: launch latency and per-block serial work dominate, not bandwidth"},
            "e2e": None, "gpu_launches": launches * K, "clocks": clocks, "cpu_baseline": None, "whole_program": prog}
    print(json.dumps(line))
    tape.close()


def bench_one_tape(args, rank, world, local_rank, W, K):
    """--split one-tape: ONE reel of the config-2 size, time-sharded over the N GPUs of the job: the cuts lie at the
    centres of inter-block gaps (shard.plan), every rank scans its own shard (plus the pre-roll rows a detector and the deskew FIFO
    need), and the results -- unit tables, proof data and events -- are gathered on rank 0 over NCCL.  STRONG scaling: the work is
    fixed, `value` = the reel's track-samples / the slowest rank's time including the gather."""
    import torch
    import torch.distributed as dist
    from readtape_b200 import shard
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    lib = abi.load_product()
    tile = synth.nrzi_tile()
    T = tile.shape[0]
    rows_total = args.rows
    hdr = synth.nrzi_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[0], hdr.bpi, hdr.ips)
    # the reel's quiet gaps: those of the super-tile, repeated (what shard.quiet_gaps finds on any stretch of it)
    lsb = hdr.maxvolts / 32767.0
    gaps_tile = shard.quiet_gaps(tile, thr_lsb=int(0.15 / lsb), min_gap_rows=2000)
    ntiles = (rows_total + T - 1) // T
    gaps = [(s + k * T, e + k * T) for k in range(ntiles) for s, e in gaps_tile if e + k * T <= rows_total]
    shards = shard.plan(rows_total, gaps, world)
    a, b = shards[rank]
    PRE = 50 + 50 + 9 + 3                                           # MAXSKEWSAMP + PKWW_MAX_WIDTH + ntrks rows in front of the cut
    lo = max(0, a - PRE) // 2048 * 2048                              # whole ingest tiles
    n = b - lo
    dev = torch.empty((n, 9), dtype=torch.int16, device="cuda")
    tile_t = torch.from_numpy(tile).cuda()
    at = 0
    while at < n:                                                   # rows [lo, b) of the periodic reel
        ph = (lo + at) % T
        m = min(T - ph, n - at)
        dev[at:at + m] = tile_t[ph:ph + m]
        at += m
    torch.cuda.synchronize()
    tape = lib.open(shard.sub_desc(desc, lo), device=local_rank)
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    recv = None

    def step():
        nonlocal recv
        tape.clear()
        tape.attach_device(dev.data_ptr(), n)
        bulk = tape.bulk_scan([cfg])
        st = bulk.stats()
        nbytes = bulk.results_size()
        img = torch.empty((nbytes + 7) // 8 * 8, dtype=torch.uint8, device="cuda")
        bulk.results_to_device(img.data_ptr(), img.numel())
        bulk.free()
        if world > 1:                                               # the result gather: sizes, then the images
            mine = torch.tensor([img.numel()], dtype=torch.int64, device="cuda")
            dist.all_gather(sizes, mine)
            mx = int(max(int(s.item()) for s in sizes))
            pad = torch.zeros(mx, dtype=torch.uint8, device="cuda"); pad[:img.numel()] = img
            if rank == 0:
                if recv is None or recv[0].numel() != mx:
                    recv = [torch.empty(mx, dtype=torch.uint8, device="cuda") for _ in range(world)]
                dist.gather(pad, recv, dst=0)
            else:
                dist.gather(pad, None, dst=0)
        return st, img.numel()

    for _ in range(W):
        step()
    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K):
        st, nbytes = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    clocks = sampler.stop()
    elapsed = t1 - t0
    ev = torch.tensor([int(st.events)], dtype=torch.int64, device="cuda")
    by = torch.tensor([int(nbytes)], dtype=torch.int64, device="cuda")
    if world > 1:
        tt = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); elapsed = float(tt.item())
        dist.all_reduce(ev); dist.all_reduce(by)
    if rank == 0:
        peak, peak_src = measured_peak()
        tsamp = rows_total * 9
        alg = 2.0 * tsamp + 32.0 * int(ev.item())
        line = {"metric": METRIC, "value": tsamp * K / elapsed, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * elapsed / K,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(rows_total).replace(", per GPU", "") + f", ONE reel time-sharded over {world} GPU(s) at inter-block gaps",
                           "split": "one-tape", "shards": [[int(x), int(y)] for x, y in shards], "pre_roll_rows": PRE,
                           "l2": "inputs far larger than the 126 MB L2", "events_on_the_reel": int(ev.item())},
                "roofline": {"bound": "hbm", "kernel": "whole step (ingest + unit finder + both scan passes + result gather)", "achieved": (alg + 2.0 * tsamp) / (elapsed / K) / 1e9 / world,
                             "peak": peak, "unit": "GB/s", "frac": (alg + 2.0 * tsamp) / (elapsed / K) / 1e9 / world / peak, "traffic": None, "peak_source": peak_src,
                             "note": "per GPU: (4 B per track-sample for the ingest + 2 B per track-sample + 32 B per event for the scan) / step time"},
                "e2e": None, "gpu_launches": (int(st.launches) + 2) * K, "clocks": clocks, "cpu_baseline": None,
                "result_gather": {"bytes_to_rank0_per_step": int(by.item()), "collective": "all_gather(sizes) + gather(images) over NCCL" if world > 1 else "none (1 GPU)"}}
        print(json.dumps(line))
    tape.close()
    if world > 1:
        dist.destroy_process_group()


def bench_gcr(args, rank, world, local_rank, W, K):
    """BASELINE config 4: synthetic 9-track GCR 6250 (9042 flux cells per inch, 50 IPS, 6.25 MHz) of DECODABLE blocks (synth.gcr_tile:
    preamble, marks, 7+1-byte groups with valid ECC and parity, resync bursts; the reference decodes them error-free), zero-crossing
    detector (-zeros, as all reference GCR examples), the 5 built-in GCR parameter sets (parmsets.c:106-110).  Full size = 50e9
    track-samples = 5.56e9 rows (100 GB) time-sharded over 8 GPUs; every GPU scans its shard for all 5 parameter sets in ONE
    rt_bulk_scan call ((parameter set x track x unit) jobs side by side).  Track-samples are counted once per parameter set.
    Weak scaling: `--rows` rows per GPU (default 1/8 of the full reel)."""
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    lib = abi.load_product()
    rows = args.rows if args.rows != FULL_ROWS else 5_555_555_556 // 8
    tile = synth.gcr_tile()
    T = tile.shape[0]
    hdr = synth.gcr_header()
    cfgs = [abi.make_cfg(tbin.MODE_GCR, p, hdr.bpi, hdr.ips, flags=abi.RT_F_FIND_ZEROS) for p in parmsets.GCR]
    dev = torch.empty((rows, 9), dtype=torch.int16, device="cuda")            # this rank's time shard of the reel: the tile sequence
    tile_t = torch.from_numpy(tile).cuda()
    for at in range(0, rows, T):
        n = min(T, rows - at)
        dev[at:at + n] = tile_t[:n]
    torch.cuda.synchronize()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns + rank * (rows // T * T) * hdr.tdelta_ns)
    tape = lib.open(desc, device=local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        tape.clear()
        tape.attach_device(dev.data_ptr(), rows)
        bulk = tape.bulk_scan(cfgs)
        st = bulk.stats()
        bulk.free()
        return st

    for _ in range(W):
        step()
    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.25)
    barrier(); t0 = time.perf_counter()
    for _ in range(K):
        st = step()
    barrier(); t1 = time.perf_counter()
    clocks = sampler.stop()
    elapsed = t1 - t0
    # full-scale check of parameter set 0 (outside the timed region): every tile the same events, one tile equal to the oracle's
    verified = None
    if not args.no_verify:
        tape.clear(); tape.attach_device(dev.data_ptr(), rows)
        bulk = tape.bulk_scan(cfgs[:1])
        verified = verify.verify_periodic(abi.load_oracle(), bulk, desc, cfgs[0], tile, rows)
        bulk.free()
    events = torch.tensor([int(st.events)], dtype=torch.int64, device="cuda")
    okt = torch.tensor([1 if (verified is None or verified["ok"]) else 0], dtype=torch.int64, device="cuda")
    if dist is not None:
        tt = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); elapsed = float(tt.item())
        dist.all_reduce(events); dist.all_reduce(okt, op=dist.ReduceOp.MIN)    # the result gather: counts only
    if rank == 0:
        peak, peak_src = measured_peak()
        npass = len(cfgs) * world
        value = npass * rows * 9 * K / elapsed
        ms_scan = float(st.ms_scan)
        alg = 2.0 * rows * 9 * len(cfgs) + 32.0 * int(st.events)
        if verified is not None:
            verified["ok_all_ranks"] = bool(okt.item())
        line = {"metric": "track-samples/s, 9-track 6.25 MHz GCR 6250 TBIN scan (zero-crossing detector), 5 parameter sets", "value": value, "unit": UNIT,
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"synthetic 9-track GCR 6250 (9042 fci, 50 IPS, 6.25 MHz) TBIN of decodable 4095-byte blocks, {rows} rows ({rows * 18 / 1e9:.1f} GB) "
                                       f"per GPU (time shard), x {len(cfgs)} parameter sets = {npass} scan passes over {world} GPU(s), -zeros",
                           "l2": "inputs far larger than the 126 MB L2", "units_per_shard": int(st.units), "events_per_gpu_step": int(st.events),
                           "tile_sha256": synth.tile_sha256(tile)[:16], "verified": verified},
                "roofline": {"bound": "hbm", "kernel": "k_units_zc (zero-crossing fast path, one lane per (parameter set, unit, track))", "achieved": alg / (ms_scan * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": alg / (ms_scan * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "ms_per_launch": ms_scan,
                             "note": "the 5 parameter sets' kernels run side by side; ms is the span from the first launch to the last completion"},
                "e2e": None, "gpu_launches": int(st.launches) * K + 2 * K, "clocks": clocks, "cpu_baseline": None,
                "result_gather": {"passes": npass, "events": int(events.item())}}
        print(json.dumps(line))
    tape.close()
    if dist is not None:
        dist.destroy_process_group()
    if not bool(okt.item()):
        sys.exit(1)


def bench_csv(args, K, W):
    """the CSV ingest.  The reference's csvtbin (src/csvtbin.c) turns the logic analyser's text export into the TBIN
    readtape reads, one fgets + ntrks scanfast_float per line on one core.  Workload: the text of the synthetic 9-track NRZI tape
    (bench tile, "%12.8f, " + 9 x "%9.5f, " per line = 114 bytes per sample, the layout csvtbin's own TBIN -> CSV direction writes),
    `--rows` lines (default 16 Mi = 1.9 GB of text).  value: conversion with the text resident on the device (k_csv_parse, timed by
    the library's CUDA events on its stream); e2e: host text -> rt_csv_open (upload + line index) -> csv_preread's maximum ->
    rt_csv_convert -> int16 rows back in pinned host memory.  cpu_baseline: the unmodified csvtbin_ref on a file of the first 2 Mi
    lines (its preread pass + conversion + .tbin written to /dev/shm)."""
    import torch
    from readtape_b200 import csvtbin
    lib = abi.load_product()
    nlines = args.rows if args.rows != FULL_ROWS else 16 << 20
    tile = synth.nrzi_tile()
    maxvolts = synth.nrzi_header().maxvolts
    ttext = synth.csv_from_rows(tile, maxvolts, 0, 1280)
    head_len = int(np.flatnonzero(ttext == 10)[1]) + 1
    body = ttext[head_len:]
    per = tile.shape[0]
    reps = (nlines + per - 1) // per
    text = np.concatenate([ttext[:head_len]] + [body] * reps)[:head_len + nlines * (len(body) // per)]
    line_bytes = len(body) // per
    ntrks = 9
    out = torch.empty((nlines, ntrks), dtype=torch.int16).pin_memory()

    def e2e_step():
        with lib.csv_open(text) as c:
            pre = csvtbin.preread(c, ntrks)
            cfg = abi.make_csv_cfg(ntrks, float(pre.maxvolts))
            st = abi.CsvStats()
            lib.check(lib.L.rt_csv_convert(c.h, ctypes.byref(cfg), 2, nlines, out.data_ptr(), None, ctypes.byref(st)))
        return pre, st

    pre, st = e2e_step()
    # parity at full size: every line of the text is a line of the tile, whose conversion the CPU oracle gives
    ora = abi.load_oracle()
    with ora.csv_open(ttext) as c:
        tile_rows, _ = c.convert(abi.make_csv_cfg(ntrks, float(pre.maxvolts)), 2, per)
    got = out.numpy()
    ok = all(np.array_equal(got[a:a + per], tile_rows[:min(per, nlines - a)]) for a in range(0, nlines, per))
    for _ in range(W):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / K
    # resident: the text stays on the device, the rows stay on the device
    sampler = ClockSampler(0)
    with lib.csv_open(text) as c:
        cfg = abi.make_csv_cfg(ntrks, float(pre.maxvolts))
        ms = []
        for _ in range(W):
            c.convert(cfg, 2, nlines, want_rows=False)
        sampler.start(); time.sleep(0.25)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(K):
            _, s = c.convert(cfg, 2, nlines, want_rows=False)
            ms.append(float(s.ms_convert))
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    clocks = sampler.stop()
    ms_k = float(np.mean(ms))
    peak, peak_src = measured_peak()
    alg = float(nlines) * (line_bytes + 2 * ntrks)
    traffic = None
    try:                                                    # dram bytes per line from the committed ncu capture of this kernel
        with open(os.path.join(ROOT, "profiles", "summary_r02.json")) as fh:
            k = json.load(fh)["k_csv_parse"]
        traffic = (k["dram_read_bytes"] + k["dram_write_bytes"]) / k["lines"] * nlines
    except Exception:
        pass
    cpu = None
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "csvtbin_ref")
    if not args.no_cpu and os.path.exists(ref_tool):
        wd = tempfile.mkdtemp(prefix="rtcsv_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            n_ref = min(nlines, 2 << 20)
            text[:head_len + n_ref * line_bytes].tofile(os.path.join(wd, "sample.csv"))
            t0 = time.perf_counter()
            subprocess.run([ref_tool, "-ntrks=9", "-nrzi", "sample"], cwd=wd, check=True, capture_output=True)
            ref_s = time.perf_counter() - t0
            cpu = {"value": n_ref * ntrks / ref_s, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"csvtbin_ref (unmodified csvtbin 1.12) on the first {n_ref} lines ({n_ref * line_bytes / 1e6:.0f} MB of text in /dev/shm): preread + conversion + .tbin written, {ref_s:.2f} s"}
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    line = {"metric": "track-samples/s, CSV text -> TBIN int16 rows (csvtbin's conversion), 9-track capture export", "value": nlines * ntrks / (ms_k * 1e-3), "unit": UNIT,
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms_k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"text export of the synthetic 9-track NRZI tape: {nlines} lines x {line_bytes} bytes = {nlines * line_bytes / 1e9:.2f} GB of CSV -> {nlines * ntrks * 2 / 1e9:.2f} GB of int16 rows",
                       "l2": "inputs far larger than the 126 MB L2", "verified": {"ok": bool(ok), "how": "every tile of the output equals the CPU oracle's conversion of the tile's text (oracle/csv_oracle.c)"},
                       "wall_ms_per_resident_step": 1e3 * dt, "too_big": int(st.too_big), "too_small": int(st.too_small)},
            "roofline": {"bound": "hbm", "kernel": "k_csv_parse (one thread per line, text staged in shared memory by cp.async.bulk)", "achieved": alg / (ms_k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms_k * 1e-3) / 1e9 / peak, "traffic": traffic, "algorithmic": alg, "peak_source": peak_src, "ms_per_launch": ms_k,
                         "algorithmic_bytes": f"{line_bytes} text bytes read + {2 * ntrks} row bytes written per line",
                         "note": "issue-bound, not memory-bound: ~36 instructions per character at 88% issue-slot use (profiles/k_csv_parse_r02_raw.csv)"},
            "e2e": {"value": nlines * ntrks / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(text.size), "d2h_bytes_per_step": int(nlines * ntrks * 2), "ms_per_step": 1e3 * e2e_s,
                    "what": "rt_csv_open (pageable host text -> device, line index) + rt_csv_max_abs over the first 999,999 lines + rt_csv_convert + rows to pinned host memory"},
            "gpu_launches": K, "clocks": clocks, "cpu_baseline": cpu}
    print(json.dumps(line))
    if not ok:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.environ.get("RT_BENCH_ROWS", FULL_ROWS)))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="nrzi", choices=["nrzi", "gcr", "pe", "ww", "csv"],
                    help="nrzi: BASELINE config 2 (the headline line); gcr: config 4, GCR-density tape x 5 parameter sets sharded over the GPUs; "
                         "pe: config 3, examples/9trk_PE with 8 parameter sets fanned out on one GPU; ww: config 5, the Whirlwind reel on the exact stateful scan")
    ap.add_argument("--split", default="replicas", choices=["replicas", "one-tape"],
                    help="N > 1: replicas = every GPU its own reel of the config-2 size (weak scaling, the default); one-tape = ONE reel "
                         "time-sharded over the GPUs at inter-block gaps, results gathered on rank 0 over NCCL (strong scaling)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-product", action="store_true", help="skip the whole-program run (TBIN file -> .tap through readtape_b200)")
    ap.add_argument("--no-verify", action="store_true", help="skip the full-scale check of the scan's events against the oracle (outside the timed regions)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = args.steps

    if args.workload == "gcr":
        return bench_gcr(args, rank, world, local_rank, W, K)
    if args.split == "one-tape" and args.impl == "b200":
        return bench_one_tape(args, rank, world, local_rank, W, K)
    if args.workload in ("pe", "ww"):
        if rank == 0:
            bench_capture(args, K, W)
        return
    if args.workload == "csv":
        if rank == 0:
            bench_csv(args, K, W)
        return
    tile = synth.nrzi_tile()
    nproc = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        work = tempfile.mkdtemp(prefix="rtref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            r = reference_run(tile, 2, nproc, K, W, work)
        finally:
            shutil.rmtree(work, ignore_errors=True)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/readtape_ref not built (run make -C oracle ref in the build container)"}))
            return
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": workload_name(args.rows), "note": "each step decodes a bounded sample of that tape"},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    numa = bind_memory_near_gpu(local_rank)
    lib = abi.load_product()                      # raises if the CUDA library is missing: no fallback
    rows = args.rows
    T = tile.shape[0]
    hdr = synth.nrzi_header()
    desc = abi.make_desc(9, hdr.maxvolts, hdr.tdelta_ns, hdr.tstart_ns)
    cfg = abi.make_cfg(tbin.MODE_NRZI, parmsets.NRZI[0], hdr.bpi, hdr.ips)

    # ---- device-resident input: the super-tile repeated (seamless by construction) ----
    dev = torch.empty((rows, 9), dtype=torch.int16, device="cuda")
    tile_t = torch.from_numpy(tile).cuda()
    for at in range(0, rows, T):
        n = min(T, rows - at)
        dev[at:at + n] = tile_t[:n]
    torch.cuda.synchronize()
    tape = lib.open(desc, device=local_rank)
    tape.prepare(cfg)                             # rt_prepare: the mask kernel runs beside the ingest kernel, chunk by chunk (RT_FUSED_MASKS=0 turns it off, =1 is the fused kernel of DESIGN.md 6b)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        tape.clear()
        tape.attach_device(dev.data_ptr(), rows)
        bulk = tape.bulk_scan([cfg])
        st = bulk.stats()
        bulk.free()
        return st

    for _ in range(W):
        st = step_resident()
    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.25)
    stats = []
    barrier(); t0 = time.perf_counter()
    for _ in range(K):
        stats.append(step_resident())
    barrier(); t1 = time.perf_counter()
    clocks = sampler.stop()
    elapsed = t1 - t0
    if dist is not None:
        tt = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
    tsamp = rows * 9
    value = world * tsamp * K / elapsed
    ms_scan = float(np.mean([s.ms_scan for s in stats])); ms_units = float(np.mean([s.ms_units for s in stats]))
    ms_ingest = float(np.mean([s.ms_preprocess for s in stats]))
    events = int(stats[-1].events); units = int(stats[-1].units)
    launches_per_step = int(stats[-1].launches) + int(stats[-1].launches_ingest)   # the scan's kernels + the ingest (and mask) kernels enqueued since rt_clear

    # ---- e2e: host buffers through the C-ABI ----
    e2e = None
    if not args.no_e2e:
        # the pinned host copy of the tape is per rank: never ask for more than a share of what the box has free
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        except Exception:
            avail = None
        rows_full = rows
        if avail is not None and rows * 18 * 1.15 > 0.6 * avail / world:
            rows = max(T, int(0.6 * avail / world / 1.15 / 18) // T * T)      # a shorter host tape (whole super-tiles) rather than an OOM
        nbytes = rows * 18
        hptr = lib.L.rt_host_alloc(nbytes)
        if not hptr:
            raise RuntimeError("rt_host_alloc failed for the pinned host tape")
        hbuf = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_int16)), shape=(rows, 9))
        for at in range(0, rows, T):
            n = min(T, rows - at)
            hbuf[at:at + n] = tile[:n]

        def step_e2e():
            bulk = tape.bulk_scan_host(hptr, rows, cfg)           # clear + H2D + ingest + scan + D2H, overlapped
            st = bulk.stats()
            hit = bulk.lookup(0, 0)                       # the user-facing read of the result
            assert hit is not None
            bulk.free()
            return st

        for _ in range(max(1, W // 2)):
            step_e2e()
        ke = max(1, K)
        barrier(); t0 = time.perf_counter()
        for _ in range(ke):
            ste = step_e2e()
        torch.cuda.synchronize(); el_local = time.perf_counter() - t0
        barrier(); t1 = time.perf_counter()
        el = t1 - t0
        if dist is not None:
            tt = torch.tensor([el], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = float(tt.item())
        e2e = {"value": world * rows * 9 * ke / el, "unit": UNIT, "h2d_bytes_per_step": int(nbytes), "rows_per_gpu": int(rows),
               "d2h_bytes_per_step": int(ste.d2h_bytes), "ms_per_step": 1e3 * el / ke, "segments_streamed": int(ste.pad),
               "h2d_GBps_this_rank": nbytes * ke / el_local / 1e9, "host_memory": numa,
               "api": "rt_bulk_scan_host (pinned host rows -> events + proof data in pinned host memory) + rt_bulk_lookup"}
        lib.L.rt_host_free(hptr)
        rows = rows_full

    # ---- full-scale verification (outside the timed regions): every tile of the periodic tape must carry the same events, every
    #      event time must be the reference's expression bit for bit, and one tile must equal the CPU oracle's events ----
    verified = None
    if not args.no_verify:
        tape.clear()
        tape.attach_device(dev.data_ptr(), rows)
        bulk = tape.bulk_scan([cfg])
        verified = verify.verify_periodic(abi.load_oracle(), bulk, desc, cfg, tile, rows)
        bulk.free()
        if not verified["ok"]:
            print(f"rank {rank}: VERIFICATION FAILED: {verified}", file=sys.stderr)
        if dist is not None:
            okt = torch.tensor([1 if verified["ok"] else 0], dtype=torch.int64, device="cuda")
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            verified["ok_all_ranks"] = bool(okt.item())

    # ---- result gather over NCCL: per-rank event counts (the only inter-GPU traffic of this path) ----
    all_events = [events]
    if dist is not None:
        g = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([events], dtype=torch.int64, device="cuda"))
        all_events = [int(x.item()) for x in g]

    # ---- the product end to end (rank 0; the reel is split over all N GPUs of the job) ----
    product = None
    if not args.no_product and rows == FULL_ROWS:
        if dist is not None:
            dist.barrier()
        if rank == 0:
            work = tempfile.mkdtemp(prefix="rtprod_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            try:
                product = product_run(tile, work, world, nproc)
            except Exception as exc:                                  # the headline numbers above stand on their own
                product = {"error": repr(exc)}
            finally:
                shutil.rmtree(work, ignore_errors=True)
        if dist is not None:
            dist.barrier()

    # ---- CPU baseline on rank 0, N=1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        work = tempfile.mkdtemp(prefix="rtcpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            r = reference_run(tile, 2, nproc, 1, 0, work)
        finally:
            shutil.rmtree(work, ignore_errors=True)
        if r is not None:
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    if rank == 0:
        peak, peak_src = measured_peak()
        ms_masks = float(np.mean([s.ms_masks for s in stats]))
        ms_records = float(np.mean([s.ms_records for s in stats]))
        force = os.environ.get("RT_SCAN")
        two_pass = ms_masks > 0 or bool(stats[-1].two_pass)
        fused = bool(stats[-1].masks_fused)
        alg_bytes = 2.0 * tsamp + 32.0 * events                # 2 B read per track-sample + event bytes
        ingest_bytes = 4.0 * tsamp                              # K1: 2 B read + 2 B written per track-sample
        # per-kernel algorithmic bytes: phase A reads every sample once and writes 3 bits per track-sample;
        # phase B reads those 3 bits and writes the events
        kernels = {"k_ingest_tma": (ms_ingest, ingest_bytes)}
        overlapped = fused and os.environ.get("RT_FUSED_MASKS", "2") == "2"
        if overlapped:                                          # ingest and mask kernels side by side on two streams, chunk by chunk: the span of both
            kernels = {"k_ingest_tma || k_peak_masks": (ms_ingest, 6.375 * tsamp)}
        elif fused:                                             # phase A runs inside the ingest kernel: 2 B read + 2 B planes + 0.375 B bit planes written per track-sample
            kernels = {"k_ingest_masks_tma": (ms_ingest, 4.375 * tsamp)}
        if two_pass and not fused:
            kernels["k_peak_masks"] = (ms_masks, 2.375 * tsamp)
        if two_pass:
            if ms_records > 0.05:
                # phase B1 reads the candidate plane and, per candidate, its window, and writes a 24-byte record; phase B2 reads the
                # records and writes the events (the candidate count is ~2 per event; counted as 2 here)
                kernels["k_cand_records"] = (ms_records, 0.125 * tsamp + 2 * events * (26.0 + 24.0))
                kernels["k_units_sparse"] = (ms_scan - ms_masks - ms_records, 2 * events * 24.0 + 32.0 * events)
            else:
                kernels["k_units_sparse"] = (ms_scan - ms_masks, 0.375 * tsamp + 32.0 * events)
        else:
            kernels["k_units_scan (generic)" if force == "generic" else "k_units_fast"] = (ms_scan, alg_bytes)
        ingest_name = "k_ingest_tma || k_peak_masks" if overlapped else "k_ingest_masks_tma" if fused else "k_ingest_tma"
        scan_kernel = max((k for k in kernels if k != ingest_name), key=lambda k: kernels[k][0])
        if kernels[ingest_name][0] > kernels[scan_kernel][0]:
            scan_kernel = ingest_name
        dom_ms, dom_bytes = kernels[scan_kernel]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None                                            # DRAM bytes of that kernel per launch, from the committed ncu capture
        try:
            with open(os.path.join(ROOT, "profiles", "summary_r02.json")) as fh:
                prof = json.load(fh).get(scan_kernel)
            if prof and prof["rows"] == rows:
                traffic = prof["dram_read_bytes"] + prof["dram_write_bytes"]
        except Exception:
            pass
        others = {k: {"ms": v[0], "algorithmic_bytes": v[1], "achieved_GBps": v[1] / (v[0] * 1e-3) / 1e9 if v[0] else None,
                      "frac": (v[1] / (v[0] * 1e-3) / 1e9 / peak) if v[0] else None} for k, v in kernels.items() if k != scan_kernel}
        others["unit_finder(5 kernels)"] = {"ms": ms_units}
        if overlapped:     # phase A's time is inside the ingest || mask pair: the scan as a whole cannot be timed apart, the step can
            ms_all = ms_ingest + ms_units + ms_scan
            others["step_total"] = {"ms": ms_all, "algorithmic_bytes": alg_bytes, "achieved_GBps": alg_bytes / (ms_all * 1e-3) / 1e9,
                                    "frac": alg_bytes / (ms_all * 1e-3) / 1e9 / peak,
                                    "note": "every kernel of the step (ingest, masks, unit finder, sparse scan) against SURVEY 8(d): 2 B per track-sample + 32 B per event"}
        else:
            others["scan_total"] = {"ms": ms_scan, "algorithmic_bytes": alg_bytes, "achieved_GBps": alg_bytes / (ms_scan * 1e-3) / 1e9,
                                    "frac": alg_bytes / (ms_scan * 1e-3) / 1e9 / peak,
                                    "note": "all scan kernels of the step against 2 B per track-sample + 32 B per event"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * elapsed / K,
            "device_ms_per_step": ms_ingest + ms_units + ms_scan,      # CUDA events on the library's own stream, summed over the step's kernels
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(rows), "l2": "inputs (rows*18 B per step) are far larger than the 126 MB L2",
                       "units_per_tape": units, "events_per_tape": events,
                       "rows_walked_one_by_one_frac": int(stats[-1].rows_scanned) / float(tsamp), "super_tile_sha256": synth.tile_sha256(tile)[:16],
                       "parallelism": f"{world} x independent tapes" if world > 1 else "1 GPU", "verified": verified},
            "roofline": {"bound": "hbm", "kernel": scan_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms, "other_kernels": others},
            "e2e": e2e, "gpu_launches": launches_per_step * K, "clocks": clocks,
            "cpu_baseline": cpu, "result_gather": {"events_per_rank": all_events}, "product": product,
        }
        print(json.dumps(line))
    tape.close()
    if dist is not None:
        dist.destroy_process_group()
    if verified is not None and not verified.get("ok_all_ranks", verified["ok"]):
        sys.exit(1)                                   # a fast scan whose events differ from the reference's is not a result


if __name__ == "__main__":
    main()
