#!/usr/bin/env python3
"""bench.py — sentences/sec of the batched Viterbi tokenizer on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on host cores

A "step" = one pass of the hot path over one batch of synthetic sentences.  At N=1 the workload is
BASELINE.json configs[2]: a unidic-cwj-3.1.1-shaped dictionary (876 803 words, 15 626 x 15 388 i16
connection matrix = 459 MiB) and a batch of 1 M ~40-character sentences.  No real dictionary exists
in this environment, so both are seeded synthetic data (vibrato_b200/synth.py) — `"data": "synthetic"`.
With N>1 every rank processes its own 1 M-sentence shard (weak scaling, no data-path collective);
rank 0 packs the dictionary image once and NCCL-broadcasts it to the other ranks.

`value`  = sentences/s with the batch already resident in HBM (device-resident C-ABI entry point),
           timed with CUDA events on the launching stream, max over ranks.
`e2e`    = the same metric through vbt_tokenize_batch with pinned HOST buffers: host->device copy
           of the sentences and device->host copy of the token records inside the timed region.
`roofline` is for the dominant kernel (k_viterbi): algorithmic bytes 2*E + 20*N (E
           connection-cost lookups of 2 B, N lattice nodes of 20 B) over its CUDA-event duration.
`cpu_baseline` = the oracle (a C restatement of vibrato's Rust path; the Rust toolchain is absent)
           timed with the reference's benchmark protocol body on one host thread, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DICT_NAME = os.environ.get("VBT_BENCH_DICT", "synth-unidic")
BATCH = int(os.environ.get("VBT_BENCH_BATCH", "1000000"))
CPU_SAMPLE = int(os.environ.get("VBT_BENCH_CPU_SAMPLE", "200000"))
METRIC = "sentences/sec (unidic-cwj-3.1.1, batch 1M) at 1/2/4/8 B200 vs ref CPU"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries (NCCL prints its version banner) may write to stdout; the contract is ONE JSON line there.
# Everything written to fd 1 during the run is diverted to stderr, the result goes to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.windows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark(self):
        """Opens / closes a timed window; only samples inside windows are reported."""
        self.windows.append(time.time())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        wins = list(zip(self.windows[0::2], self.windows[1::2]))
        inside = [ln for (t, ln) in self.lines if any(a - 0.05 <= t <= b + 0.15 for a, b in wins)]
        if not inside:  # very short runs: fall back to everything sampled since start()
            inside = [ln for (_, ln) in self.lines]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(rank, need_matrix):
    from vibrato_b200 import synth
    t = time.time()
    sd = synth.make_dictionary(DICT_NAME, with_matrix=need_matrix)
    log(f"[rank {rank}] synthetic dictionary {DICT_NAME}: {time.time() - t:.1f}s")
    t = time.time()
    utf8, off = synth.make_corpus(sd, BATCH, seed=20260923 + 2 + 1000 * rank)
    log(f"[rank {rank}] corpus {BATCH} sentences, {len(utf8) / 1e6:.1f} MB: {time.time() - t:.1f}s")
    return sd, utf8, off


def workload_config(extra=None):
    from vibrato_b200 import synth
    nw, nr, nl, _, _ = synth.SHAPES[DICT_NAME]
    cfg = {
        "workload": f"{DICT_NAME}: synthetic stand-in for unidic-cwj-3.1.1 ({nw} words, {nl}x{nr} i16 connection "
                    f"matrix = {nl * nr * 2 / 2**20:.0f} MiB), batch {BATCH} synthetic ~40-char JA sentences per GPU "
                    "(BASELINE.json configs[2])",
        "dictionary": "synthetic (no real unidic/ipadic in this environment)",
        "batch_per_gpu": BATCH,
        "sentence_len_chars": "round(N(40,8^2)) clipped to [8,120]",
        "cache": "inputs_exceed_l2 (sentences + lattice workspace per step are GBs; L2 is 126 MB)",
        "parallelism": "shard-over-sentences, replicated dictionary",
    }
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path = the oracle (C restatement; the Rust
    crate cannot be built here), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import vibrato_oracle as vo
    sd, utf8, off = make_inputs(0, True)
    t = time.time()
    od = vo.OracleDictionary(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
    log(f"oracle dictionary build: {time.time() - t:.1f}s")
    threads = os.cpu_count() or 1
    n = min(CPU_SAMPLE, BATCH)
    sub_off = off[: n + 1]
    for _ in range(args.warmup):
        od.benchmark(utf8, sub_off, n_threads=threads, runs=1)
    secs, nwords = od.benchmark(utf8, sub_off, n_threads=threads, runs=args.steps)
    value = n * args.steps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sentences/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config({"sample": f"first {n} sentences of the batch per step"}),
        "cpu_baseline": {"value": value, "unit": "sentences/s", "cores": threads, "kind": "port",
                         "sample": f"{n} sentences x {args.steps} steps, {threads} threads; oracle = C restatement of "
                                   "vibrato 0.5.2 (Rust toolchain absent), protocol body of benchmark/src/main.rs:53-65"},
        "e2e": {"value": value, "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "n_words": int(nwords),
    }
    emit(line)


def run_ours(args, rank, world, local_rank):
    import torch
    import vibrato_b200 as vb
    from vibrato_b200._native import check, lib
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the tokenizer has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sd, utf8, off = make_inputs(rank, rank == 0)
    n_bytes = int(off[-1])

    # --- dictionary image: rank 0 packs it, everyone else receives it over NCCL ------------------
    t = time.time()
    if rank == 0:
        d = vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
        blob_h = d.pack_blob()
        size = torch.tensor([blob_h.nbytes], dtype=torch.int64, device="cuda")
        log(f"[rank 0] host dictionary + image ({blob_h.nbytes / 2**20:.0f} MiB): {time.time() - t:.1f}s")
    else:
        size = torch.zeros(1, dtype=torch.int64, device="cuda")
    if dist:
        dist.broadcast(size, 0)
    blob = torch.empty(int(size.item()), dtype=torch.uint8, device="cuda")
    if rank == 0:
        blob.copy_(torch.from_numpy(blob_h))
    if dist:
        dist.broadcast(blob, 0)  # the one collective of the path: dictionary image over NVLink
    torch.cuda.synchronize()
    h = C.c_void_p()
    check(lib().vbt_tokenizer_new_from_device_blob(blob.data_ptr(), blob.numel(), 0, 0, local_rank, C.byref(h)))
    stream = torch.cuda.current_stream()
    check(lib().vbt_tokenizer_set_stream(h, stream.cuda_stream))
    for opt in ("lanes_per_sentence", "sort_by_length", "chunk_sentences", "dual_stream"):  # developer overrides
        if os.environ.get("VBT_" + opt.upper()):
            check(lib().vbt_tokenizer_set_option(h, opt.encode(), int(os.environ["VBT_" + opt.upper()])))

    # --- inputs: pinned host copies (e2e) and device-resident copies (value) ----------------------
    h_utf8 = torch.from_numpy(utf8).pin_memory()
    h_off = torch.from_numpy(off.astype(np.int64)).pin_memory()
    d_utf8 = h_utf8.cuda()
    d_off = h_off.cuda()
    torch.cuda.synchronize()

    def step_device():
        a, b, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().vbt_tokenize_batch_device(h, d_utf8.data_ptr(), d_off.data_ptr(), BATCH, n_bytes, C.byref(a),
                                              C.byref(b), C.byref(n)))
        return n.value

    def step_host():
        r = C.c_void_p()
        check(lib().vbt_tokenize_batch(h, h_utf8.data_ptr(), h_off.data_ptr(), BATCH, C.byref(r)))
        nt = C.c_uint64()
        check(lib().vbt_result_view(r, None, None, None, C.byref(nt)))
        lib().vbt_result_free(r)
        return nt.value

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def stage_ms():
        ms = (C.c_float * 16)()
        n = C.c_int32()
        check(lib().vbt_last_stage_ms(h, ms, 16, C.byref(n)))
        return np.array([ms[i] for i in range(n.value)], dtype=np.float64)

    stage_names = lib().vbt_stage_names().decode().split(",")

    # one counted batch (outside the timed region): E, N and the whole-path B_alg of this workload
    check(lib().vbt_tokenizer_set_counting(h, 1))
    n_tokens = step_device()
    cnt = (C.c_uint64 * 10)()
    check(lib().vbt_last_counters(h, cnt))
    cnt = np.array(list(cnt), dtype=np.float64)
    check(lib().vbt_tokenizer_set_counting(h, 0))
    w = np.array([1, 4, 4, 8, 4, 6, 2, 20, 24, 0], dtype=np.float64)
    b_alg_step = float((cnt * w).sum())
    b_alg_viterbi = float(2 * cnt[6] + 20 * cnt[7])
    nl = C.c_uint64()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = np.zeros(len(stage_names))
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        stage_acc += stage_ms()
    e1.record(stream)
    barrier()
    sampler.mark()
    dev_ms = e0.elapsed_time(e1)
    check(lib().vbt_last_launch_count(h, C.byref(nl)))
    launches_per_step = nl.value

    # --- e2e: host buffers in, host tokens out ----------------------------------------------------
    for _ in range(2):
        step_host()
    barrier()
    sampler.mark()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        nt_host = step_host()
    f1.record(stream)
    barrier()
    sampler.mark()
    e2e_ms = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    assert nt_host == n_tokens

    t_dev = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t_dev[0]), float(t_dev[1])

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        st = stage_acc / args.steps
        vit_ms = float(st[stage_names.index("viterbi")])
        achieved = b_alg_viterbi / (vit_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if tj.get("dict") == DICT_NAME and tj.get("batch") == BATCH:
                traffic = tj.get("viterbi_dram_bytes_per_launch")
        except Exception:
            pass
        # CPU baseline (rank 0, N=1 only): the oracle, one thread, bounded sample
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import vibrato_oracle as vo
            t = time.time()
            od = vo.OracleDictionary(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
            n = min(CPU_SAMPLE, BATCH)
            od.benchmark(utf8, off[: n // 10 + 1], n_threads=1, runs=1)  # warm-up (benchmark/src/main.rs:69-72)
            secs, _ = od.benchmark(utf8, off[: n + 1], n_threads=1, runs=2)
            cpu = {"value": 2 * n / secs, "unit": "sentences/s", "cores": 1, "kind": "port",
                   "sample": f"first {n} sentences of the batch x 2 runs, 1 thread; oracle = C restatement of vibrato "
                             "0.5.2 (Rust toolchain absent), body of benchmark/src/main.rs:53-65"}
            log(f"cpu baseline: {cpu['value']:.0f} sentences/s ({time.time() - t:.1f}s)")
        total = BATCH * world
        h2d = int(n_bytes + (BATCH + 1) * 8)
        d2h = int((BATCH + 1) * 8 + n_tokens * 24)
        line = {
            "metric": METRIC, "value": total * args.steps / (dev_ms * 1e-3), "unit": "sentences/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(),
            "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": "sentences/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "kernel": "k_viterbi", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_alg_viterbi, "kernel_ms": vit_ms,
                         "whole_step_algorithmic_bytes": b_alg_step,
                         "whole_step_achieved_GBs": b_alg_step / (dev_ms / args.steps * 1e-3) / 1e9},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "stage_ms": dict(zip(stage_names, [round(float(x), 4) for x in st])),
            "tokens_per_step": int(n_tokens),
            "counters_per_sentence": dict(zip("U C M T P W E N K walks".split(), [round(float(x) / BATCH, 2) for x in cnt])),
        }
        emit(line)
    lib().vbt_tokenizer_free(h)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
