#!/usr/bin/env python
"""bench.py -- mel frames/sec of the Tacotron 2 hot path (BASELINE.json configs[1]).

One "step" = one pass of the hot path over one batch: Tacotron2.inference on (B=64 per GPU,
T_text=150) synthetic text, exactly 800 decoder frames per row (gate_threshold = 1.0 so the stop
gate never fires, max_decoder_steps = 800) -> encoder, 800-step persistent
decoder, postnet.  51,200 mel frames per GPU per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  value : frames/s with the text ids already resident in HBM (device tensors through the nn.Module API)
  e2e   : frames/s through the C-ABI t2_infer_host with HOST buffers (pinned text in, mel_postnet out)
  roofline     : the persistent decoder kernel, algorithmic FLOPs (38,350,592 per frame) / CUDA-event time
  cpu_baseline : the oracle port (oracle/tacotron2_oracle.py, torch CPU, all host threads) on a bounded sample
  --impl reference : the same metric from the CPU oracle port alone (the reference is pure Python and
                     /root/reference does not exist on the GPU box; DESIGN.md "reference arm")
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, T_TEXT, T_MEL = 64, 150, 800
FLOP_PER_FRAME = 2 * (18167296 + 6720 * T_TEXT)          # 38,350,592 @ T_enc=150
STREAM_BYTES_PER_STEP = 97.3e6                           # fp32 weights + memory + processed memory


def synth_weights(seed=1234):
    from tests.common import synth_state_dict
    return synth_state_dict(seed, gate_bias=0.0, scale=1.0)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.rows)}


def _best_threads(fn, candidates):
    """Runs fn() under each thread count and returns (best_seconds, best_threads)."""
    best = None
    for th in candidates:
        torch.set_num_threads(th)
        fn()                                  # warm-up at this thread count
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, th)
    return best


def cpu_port_sample(threads, dec_steps=60):
    """Times the oracle port (the reference's algorithm in plain torch CPU ops) on the host cores on a
    bounded sample of the SAME workload: encoder + `dec_steps` of the 800 decoder steps + postnet at B=64,
    T_text=150; the decoder part is extrapolated linearly to 800 steps (every step does identical work).
    Each component runs at the thread count (<= all host threads) that is fastest for it -- small
    recurrent GEMMs are slower with 100+ threads than with 16."""
    from oracle import tacotron2_oracle as O
    from tests.common import keep_mask, rand_text
    cands = sorted({t for t in (8, 16, 32, 64, threads) if t <= threads})
    sd = synth_weights()
    text = rand_text(B_PER_GPU, T_TEXT, 1)
    keep = keep_mask((dec_steps + 3, 2, B_PER_GPU, 256), 0.5, 2)
    with torch.no_grad():
        emb = sd["embedding.weight"][text].transpose(1, 2)
        # encoder: tune on a 1/5 slice of the sequence, then time the full one
        _, th_enc = _best_threads(lambda: O.encoder(sd, emb[:, :, :30]), cands)
        torch.set_num_threads(th_enc)
        t0 = time.perf_counter()
        memory = O.encoder(sd, emb)
        t_enc = time.perf_counter() - t0
        st0 = O.init_decoder_state(sd, memory)

        def steps(n, st=None):
            st = st or {k: v.clone() for k, v in st0.items()}
            x = memory.new_zeros(B_PER_GPU, 80)
            ts = []
            for t in range(n):
                t0 = time.perf_counter()
                px = O.prenet(sd, x, keep[t, 0], keep[t, 1])
                x, _, _ = O.decode_step(sd, st, memory, px)
                ts.append(time.perf_counter() - t0)
            return ts
        _, th_dec = _best_threads(lambda: steps(6), cands)
        torch.set_num_threads(th_dec)
        times = sorted(steps(dec_steps + 3)[3:])
        step = times[len(times) // 2]
        mel = torch.randn(B_PER_GPU, 80, T_MEL)
        _, th_post = _best_threads(lambda: O.postnet(sd, mel[:, :, :100]), cands)
        torch.set_num_threads(th_post)
        t0 = time.perf_counter()
        O.postnet(sd, mel)
        t_post = time.perf_counter() - t0
    total = t_enc + T_MEL * step + t_post
    # cores = the threads actually used (the largest per-component choice; more threads made every component slower)
    return {"value": B_PER_GPU * T_MEL / total, "unit": "mel frames/s", "cores": max(th_enc, th_dec, th_post), "kind": "port",
            "sample": "oracle port, B=64 T_text=150, fp32, %d host threads available: encoder %.3f s (%d thr) + median of "
                      "%d decoder steps %.1f us/step x800 (%d thr) + postnet T_mel=800 %.3f s (%d thr)"
                      % (threads, t_enc, th_enc, dec_steps, step * 1e6, th_dec, t_post, th_post),
            "decoder_step_us": step * 1e6}, total


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        cb, total = cpu_port_sample(threads, dec_steps=20 if i < args.warmup else 40)
        if i >= args.warmup:
            vals.append((cb, total))
    cb = vals[len(vals) // 2][0]
    ms = sum(v[1] for v in vals) / len(vals) * 1e3
    line = {"impl": "reference", "metric": "mel frames/sec (B=64,T_text=150)", "value": cb["value"], "unit": "mel frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Tacotron2.inference B=64 T_text=150 T_mel=800 (BASELINE.json configs[1]) on host CPU cores; "
                                   "each step = bounded sample extrapolated to 800 decoder steps"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "mel frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--decoder-impl", default="auto", choices=["auto", "stepwise", "persistent"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import tacotron2_b200 as t2
    from tacotron2_b200 import _capi
    from tests.common import rand_text
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    hp = t2.create_hparams()
    model = t2.Tacotron2(hp)
    model.load_state_dict(synth_weights())
    model = model.cuda().eval()
    model.decoder.max_decoder_steps = T_MEL
    model.decoder.gate_threshold = 1.0            # sigmoid(.) > 1.0 never fires -> exactly 800 frames per row
    eng = model._t2_engine()
    eng.impl = {"auto": _capi.IMPL_AUTO, "stepwise": _capi.IMPL_STEPWISE, "persistent": _capi.IMPL_PERSISTENT}[args.decoder_impl]
    L = _capi.lib()

    text = rand_text(B_PER_GPU, T_TEXT, 100 + rank)
    text_dev = text.cuda()
    text_host = text.clone().pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")    # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters):
        """Per-iteration CUDA-event timing (L2 flushed, untimed, between iterations)."""
        total = 0.0
        for _ in range(iters):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    import contextlib

    def step_device():
        # the reference prints "Warning! Reached max decoder steps" to stdout (model.py:446); this
        # workload reaches the cap by construction, keep stdout for the single JSON line
        with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):
            out = model.inference(text_dev)
        return out

    out_host = None

    def step_host():
        nonlocal out_host
        out_host = eng.infer_host(text_host, T_MEL, 1.0, out_host=out_host)

    # component timers (rank-local, for the roofline / breakdown)
    def decoder_only():
        with torch.no_grad():
            return eng.decoder(memory_dev, _capi.MODE_INFER, T_MEL, gate_threshold=1.0)

    with torch.no_grad():
        memory_dev = eng.encoder(text=text_dev)
    for _ in range(args.warmup):
        step_device(); step_host()
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    launches0 = L.t2_kernel_launch_count()
    ms_dev = timed(step_device, args.steps)
    launches = L.t2_kernel_launch_count() - launches0
    barrier()
    ms_e2e = timed(step_host, args.steps)
    barrier()
    ms_dec = timed(decoder_only, args.steps)
    sampler.stop_flag = True
    phase_profile = None
    try:
        prof = eng.decoder_profile()
        tot = sum(v[0] for v in prof.values()) or 1
        phase_profile = {k: {"us_per_step_cta0_60_100": [round(x / 1965.0 / T_MEL, 2) for x in v]} for k, v in prof.items()}
    except Exception as e:  # stepwise implementation has no phase profile
        phase_profile = {"unavailable": str(e)[:80]}
    n_frames = int(out_host[2][0]) * B_PER_GPU
    assert n_frames == B_PER_GPU * T_MEL, "workload did not produce 800 frames per row: %d" % n_frames
    t = torch.tensor([ms_dev, ms_e2e, ms_dec], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_dec = (float(x) for x in t.cpu())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames = B_PER_GPU * T_MEL * world * args.steps
    value = frames / (ms_dev * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    peak_tf, peak_gbs, peak_src = load_peaks()
    dec_s = ms_dec * 1e-3 / args.steps
    ach_tf = B_PER_GPU * T_MEL * FLOP_PER_FRAME / dec_s / 1e12
    ach_gbs = T_MEL * STREAM_BYTES_PER_STEP / dec_s / 1e9
    info = (torch.cuda.get_device_name(0))
    line = {
        "metric": "mel frames/sec (B=64,T_text=150)", "value": value, "unit": "mel frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (split-fp16 tensor-core operands hi+lo, fp32 accumulate and state)", "data": "synthetic",
        "config": {"workload": "Tacotron2.inference: B=64 per GPU, T_text=150, 800 decoder frames per row "
                               "(gate_threshold=1.0, max_decoder_steps=800), encoder + decoder + postnet; BASELINE.json configs[1]",
                   "global_batch": B_PER_GPU * world, "parallelism": "dp%d (batch sharded, no collective)" % world,
                   "l2": "256 MiB flush between timed iterations", "decoder_impl": args.decoder_impl, "device": info},
        "e2e": {"value": e2e, "unit": "mel frames/s", "h2d_bytes_per_step": B_PER_GPU * T_TEXT * 8,
                "d2h_bytes_per_step": B_PER_GPU * 80 * T_MEL * 4 + B_PER_GPU * 4 + 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "decoder_step_us": dec_s / T_MEL * 1e6, "decoder_ms": dec_s * 1e3,
        "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                     # dram__bytes_read+write of the persistent kernel from the ncu --set full capture in
                     # profiles/r01_decoder_final_ncu_summary.md (3.796 GB per 100-step launch), scaled to 800 steps
                     "traffic": 3.796e9 * T_MEL / 100, "kernel": "decoder (persistent kernel + processed_memory GEMM), CUDA events",
                     "peak_source": peak_src, "algorithmic_flop_per_frame": FLOP_PER_FRAME,
                     "stream_bytes": {"achieved_GBps": ach_gbs, "peak_GBps": peak_gbs, "frac": ach_gbs / peak_gbs,
                                      "bytes_per_step": STREAM_BYTES_PER_STEP}},
        "clocks": sampler.summary(),
        "decoder_phase_profile": phase_profile,
    }
    if not args.no_cpu_baseline and world == 1:
        cb, _ = cpu_port_sample(os.cpu_count() or 1, dec_steps=60)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
