#!/usr/bin/env python
"""tools/fuzz_campaign.py -- run the seeded CPU fuzz tests of tests/ over seed ranges far beyond what the suite runs.

    python tools/fuzz_campaign.py proof 100 400        # tests/test_proof_host.py, seeds 100..399
    python tools/fuzz_campaign.py ctx 100 400          # tests/test_ctx_host.py (exact scan with skip-ahead)
    python tools/fuzz_campaign.py sparse 100 400       # tests/test_sparse_host.py adversarial signals (one seed per call)
    python tools/fuzz_campaign.py generic 100 400      # tests/test_proof_generic_host.py (k_units_scan's detectors and proof data)
    python tools/fuzz_campaign.py shim-synth 0 200     # tests/test_fuzz_shim.py: the event-driven readblock() on the oracle backend beside the
    python tools/fuzz_campaign.py shim-capture 0 200   #   unmodified reference binary (4 random cases per seed): rc, outputs and logs identical
    python tools/fuzz_campaign.py sim-synth 0 200      # the same on the CPU simulation of the whole C-ABI (tests/host_fast/hostsim.cu): speculative
    python tools/fuzz_campaign.py sim-capture 0 200    #   hits, restarts, bridge scans, with units cut at random
    python tools/fuzz_campaign.py sim-dropout 0 200    #   synthetic tapes with all-track drop-outs (the real unit finder cuts inside blocks)
    python tools/fuzz_campaign.py sim-workers 0 200    #   one reel split between worker processes (RT_WORKERS) with small shares and random unit cuts
    python tools/fuzz_campaign.py oracle 100 200       # tests/test_oracle_fuzz.py (the instrumented unmodified reference)

Each seed runs in its own pytest-free call of the test function; a failing seed is printed and the campaign goes on.  TEST
INFRASTRUCTURE: everything here runs the host builds and the oracle, never the product path.
"""
import os, sys, tempfile, pathlib, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import subprocess


def main():
    which, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    from readtape_b200 import abi
    import test_fast_host as tfh
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host_fast")], stdout=subprocess.DEVNULL)
    oracle = abi.load_oracle()
    bad = []
    def host_lib():
        import test_sparse_host as tsh
        return tsh.host_lib.__wrapped__() if hasattr(tsh.host_lib, "__wrapped__") else None
    L = C.CDLL(tfh.HOST_LIB)
    L.fast_host_scan_unit.restype = C.c_int
    L.fast_host_scan_unit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg), C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    L.sparse_host_scan_unit.restype = C.c_int
    L.sparse_host_scan_unit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(abi.TapeDesc), C.POINTER(abi.ScanCfg), C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
    L.masks_host_build.restype = C.c_int
    L.masks_host_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    for seed in range(lo, hi):
        try:
            if which == "proof":
                import test_proof_host as m
                try: m.test_accepted_resets_reproduce_the_fresh_scan.__wrapped__
                except AttributeError: pass
                m.test_accepted_resets_reproduce_the_fresh_scan(range(seed, seed + 1), L, oracle)
            elif which == "ctx":
                import test_ctx_host as m
                m.test_exact_scan_with_skip_ahead_equals_oracle(range(seed, seed + 1), L, oracle)
            elif which == "sparse":
                import test_sparse_host as m
                m.test_sparse_scan_on_adversarial_signals(seed, L, oracle)
            elif which == "generic":
                import test_proof_generic_host as m
                m.test_generic_unit_scan_and_its_proof_data(range(seed, seed + 1), L, oracle)
            elif which in ("shim-synth", "shim-capture"):
                import test_fuzz_shim as m
                with tempfile.TemporaryDirectory() as d:
                    m.fuzz(m.ORACLE_SHIM, m.synthetic_case if which == "shim-synth" else m.capture_case, 100000 + seed, 4, pathlib.Path(d))
            elif which in ("sim-synth", "sim-capture", "sim-dropout"):
                import test_fuzz_shim as m, test_hostsim as hsim
                case = hsim._with_random_cuts({"sim-synth": m.synthetic_case, "sim-capture": m.capture_case, "sim-dropout": hsim.dropout_case}[which])
                with tempfile.TemporaryDirectory() as d:
                    try: m.fuzz(hsim.SIM, case, 200000 + seed, 4, pathlib.Path(d))
                    finally: [os.environ.pop(k, None) for k in hsim.SIM_KNOBS]
            elif which in ("sim-workers", "sim-workers-capture"):
                import test_hostsim as hsim
                with tempfile.TemporaryDirectory() as d:
                    hsim.worker_fuzz(hsim.SIM, 300000 + seed, 4, pathlib.Path(d), hsim.worker_capture_case if which.endswith("capture") else None)
            elif which == "oracle":
                import test_oracle_fuzz as m
                with tempfile.TemporaryDirectory() as d:
                    m.test_oracle_reproduces_the_reference_events_on_adversarial_captures(range(seed, seed + 1), oracle, pathlib.Path(d))
            else:
                raise SystemExit("unknown campaign " + which)
        except AssertionError as e:
            msg = str(e)
            # the sanity floors of the tests (enough accepted / bridged / jumped cases in a RANGE of seeds: the message is the tuple
            # of counts) do not apply to a single seed
            if msg.startswith("(") and msg.rstrip().endswith(")") and "\n" not in msg: continue
            bad.append(seed); print(f"[{which}] seed {seed} FAILED: {msg[:600]}", flush=True)
        except BaseException as e:                                  # pytest.fail raises Failed (a BaseException)
            if isinstance(e, (KeyboardInterrupt, SystemExit)): raise
            bad.append(seed); print(f"[{which}] seed {seed} FAILED: {type(e).__name__}: {str(e)[:600]}", flush=True)
        if (seed - lo) % 20 == 19: print(f"[{which}] .. seed {seed}, {len(bad)} failures so far", flush=True)
    print(f"[{which}] seeds {lo}..{hi - 1}: {len(bad)} failing seeds {bad}", flush=True)


if __name__ == "__main__":
    main()
