#!/usr/bin/env python
"""bench.py -- KBNet forward frames/s @352x1216 on N x B200 (BASELINE.json metric), with the S2D / KB rooflines.

  python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path (oracle port), rank 0 only

A "step" = one forward of the full KBNet (S2D -> encoder with 4 KB layers -> decoder -> depth head) over a batch of
8 synthetic KITTI-shaped frames per GPU (BASELINE config 3; config 5 is the same per-GPU work on 8 GPUs: weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'kbnet_fwd_frames_per_sec_352x1216'
H, W = 352, 1216
MIN_POOLS, MAX_POOLS = [5, 7, 9, 11, 13], [15, 17]
F_IMG, F_DEP, F_DEC = [48, 96, 192, 384, 384], [16, 32, 64, 128, 128], [256, 128, 128, 64, 12]
KITTI_K = [[721.5377, 0.0, 609.5593], [0.0, 721.5377, 172.854], [0.0, 0.0, 1.0]]


def synth_inputs(n, seed):
    gen = torch.Generator().manual_seed(seed)
    image = torch.rand(n, 3, H, W, generator=gen)
    mask = (torch.rand(n, 1, H, W, generator=gen) < 0.05).float()
    depth = (torch.rand(n, 1, H, W, generator=gen) * 79.0 + 1.0) * mask
    valid = (depth > 0).float()
    k = torch.tensor(KITTI_K, dtype=torch.float32).view(1, 3, 3).repeat(n, 1, 1)
    return image, depth, valid, k


def ncu_traffic(key):
    """dram bytes of one launch of the kernel, from the committed ncu capture (profiles/r01_traffic.json); None if absent"""
    try:
        return float(json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')))[key]['traffic'])
    except Exception:
        return None


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return float(p['hbm_gbs']), float(p['bf16_tflops']), float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), 'measured'
    except Exception:
        return 6650.0, 1590.0, 1400.0, 'fallback'


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe): NVML every ~5 ms when
    nvidia_ml_py is importable (a bench step is a few ms), else nvidia-smi polling."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(',') if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            bits = (('hw_slowdown', n.nvmlClocksThrottleReasonHwSlowdown), ('hw_thermal_slowdown', n.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ('sw_thermal_slowdown', n.nvmlClocksThrottleReasonSwThermalSlowdown), ('sw_power_cap', n.nvmlClocksThrottleReasonSwPowerCap))
            try:
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            except Exception:
                mx = None
            while not self.stop_flag:
                try:
                    sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                    self.rows.append([str(sm), str(mx), str(pw)] + ['Active' if rs & b else 'Not Active' for _, b in bits])
                except Exception:
                    pass
                time.sleep(0.005)
            return
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        mx = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()), default=None)
        pw = max((float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace('.', '').isdigit()), default=None)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(self.rows),
                'power_w_max': pw, 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


def cpu_forward_fn():
    """The reference's CPU path for this workload: the oracle's op-for-op restatement (same ATen CPU ops as the
    reference modules; /root/reference itself does not exist on the GPU box)."""
    from oracle import kbnet_oracle as K
    ws, we, wd = K.make_model_weights(torch.Generator().manual_seed(0), 2, 7, 8, 3, F_IMG, F_DEP, F_DEC)

    def fwd(image, depth, valid, k):
        with torch.no_grad():
            return K.model_forward(image, depth, valid, k, ws, we, wd, MIN_POOLS, MAX_POOLS, 1.5, 100.0)
    return fwd


def run_reference(args, rank):
    """--impl reference: CPU reference path, all host threads, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd = cpu_forward_fn()
    nb = args.cpu_batch
    inp = synth_inputs(nb, 1)
    for _ in range(max(args.warmup, 1) if args.warmup < 2 else 1):   # one warm-up pass is enough on the CPU
        fwd(*inp)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        fwd(*inp)
        times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    val = nb / t
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'kbnet_kitti_352x1216_full_forward', 'frames_per_step': nb, 'device': 'cpu'},
            'cpu_baseline': {'value': val, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d frames per step, oracle port of the reference PyTorch-CPU path (same ATen ops)' % nb},
            'e2e': {'value': val, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='kbnet', choices=['kbnet', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='frames per GPU per step')
    ap.add_argument('--cpu-batch', type=int, default=2, help='frames per step of the CPU reference sample')
    ap.add_argument('--conv-backend', default=None, choices=['kbnet', 'torch'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-autotune', action='store_true')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    import torch.distributed as dist
    import kbnet_b200
    import kbnet_b200.model as M
    if args.conv_backend:
        M.CONV_BACKEND = args.conv_backend
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    torch.set_grad_enabled(False)
    lib = kbnet_b200._lib.load()
    kbnet_b200.set_autotune(not args.no_autotune)     # tile plans are measured during the warm-up steps

    torch.manual_seed(0 if rank == 0 else 1234 + rank)         # only rank 0's weights survive the broadcast
    model = kbnet_b200.KBNetModel(3, 2, MIN_POOLS, MAX_POOLS, 3, 8, F_IMG, F_DEP, [0, 1, 2, 3], F_DEC, deconv_type='up',
                                  weight_initializer='xavier_normal', activation_func='leaky_relu',
                                  min_predict_depth=1.5, max_predict_depth=100.0, device=dev)
    bcast_bytes = kbnet_b200.dist.broadcast_parameters(model.parameters(), src=0) if world > 1 else 0

    nb = args.batch
    host = [t.pin_memory() for t in synth_inputs(nb, 1 + rank)]
    resident = [t.to(dev) for t in host]
    out_host = torch.empty((nb, 1, H, W), dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * 4 for t in host)
    d2h = out_host.numel() * 4

    # per-kernel CUDA events on the launching stream (torch's current stream) around the S2D call
    s2d_events = []
    s2d_mod = model.sparse_to_dense_pool
    orig_forward = s2d_mod.forward

    def timed_s2d(x, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = orig_forward(x, *a, **kw)
        e1.record()
        s2d_events.append((e0, e1))
        return y
    s2d_mod.forward = timed_s2d

    # ... and around KB block 2 (BASELINE config 2 is exactly this call at 8 frames: 48/16/48 ch @176x608 -> 96/32/96 @88x304)
    kb_events = []
    kb2 = model.encoder.calibrated_backprojection2
    kb2_forward = kb2.forward

    def timed_kb2(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = kb2_forward(*a, **kw)
        e1.record()
        kb_events.append((e0, e1))
        return y
    kb2.forward = timed_kb2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        model.forward(*resident)
    barrier()
    s2d_events.clear()
    kb_events.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = lib.kbnet_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = model.forward(*resident)
    e1.record()
    barrier()
    launches = lib.kbnet_launch_count() - n0
    ms = e0.elapsed_time(e1)
    s2d_ms = sum(a.elapsed_time(b) for a, b in s2d_events) / max(len(s2d_events), 1)
    kb_ms = sum(a.elapsed_time(b) for a, b in kb_events) / max(len(kb_events), 1)
    tma_launches = lib.kbnet_tma_engine_launch_count()

    # end-to-end through the public API: every step copies ITS inputs from pinned host memory and reads ITS result back.
    # Two device input sets + a copy stream: the H2D of step i+1 overlaps the forward of step i (the usual serving pipeline).
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    dev_in = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(in_free[slot])
            for d, h in zip(dev_in[slot], host):
                d.copy_(h, non_blocking=True)
            in_ready[slot].record(copy_stream)

    def e2e_loop(n_steps):
        for ev in in_free:
            ev.record(main_stream)
        upload(0)
        for i in range(n_steps):
            slot = i & 1
            if i + 1 < n_steps:
                upload(slot ^ 1)
            main_stream.wait_event(in_ready[slot])
            out_host.copy_(model.forward(*dev_in[slot]), non_blocking=True)
            in_free[slot].record(main_stream)

    e2e_loop(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(args.steps)
    copy_stream.synchronize()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.summary() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms, ms_e2e, s2d_ms, kb_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, s2d_ms, kb_ms = (float(v) for v in t.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = nb * world * args.steps
    value = frames / (ms * 1e-3)
    e2e_val = frames / (ms_e2e * 1e-3)
    hbm_peak, tf_peak, tf_sus, peak_kind = peaks()
    s2d_bytes = 4.0 * H * W * (2 + 8) * nb                    # 4*H*W*(C_in + n_filter) per frame
    s2d_gbs = s2d_bytes / (s2d_ms * 1e-3) / 1e9 if s2d_ms > 0 else 0.0
    # KB block 2 (fast path): image 48ch + depth 16ch full reads, fused 48ch at even pixels, three outputs
    kb_bytes = 4.0 * nb * ((48 + 16) * 176 * 608 + 48 * 88 * 304 + (96 + 32 + 96) * 88 * 304)
    kb_flops = 24.19e9 * nb / 8.0
    kb_gbs = kb_bytes / (kb_ms * 1e-3) / 1e9 if kb_ms > 0 else 0.0
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'kbnet_kitti_352x1216_full_forward', 'frames_per_gpu_per_step': nb, 'global_batch': nb * world,
                   'pools': {'min': MIN_POOLS, 'max': MAX_POOLS}, 'parallelism': 'batch-sharded x%d, weights broadcast once (%d B)' % (world, bcast_bytes),
                   'conv_backend_outside_hot_path': M.CONV_BACKEND, 'tile_autotune': not args.no_autotune,
                   'l2': 'per-step activations (>2 GB at batch 8) far exceed the 126 MB L2; no explicit flush',
                   'e2e_pipeline': 'double-buffered pinned H2D on a copy stream overlapped with the previous forward; D2H of every result'},
        'e2e': {'value': e2e_val, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d * world, 'd2h_bytes_per_step': d2h * world,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'kernel': 's2d_fused_kernel', 'bound': 'hbm', 'achieved': s2d_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': s2d_gbs / hbm_peak, 'traffic': ncu_traffic('s2d_fused_kernel') if nb == 8 else None, 'peak_kind': peak_kind,
                     'algorithmic_bytes_per_launch': s2d_bytes, 'ms_per_launch': s2d_ms,
                     'traffic_note': 'dram__bytes_read+write of one launch, ncu --set full, profiles/r01_traffic.json'},
        'roofline_kb': {'kernel': 'conv_v2_kernel<S2K3> (fused KB block 2 = BASELINE config 2)', 'bound': 'hbm', 'achieved': kb_gbs,
                        'peak': hbm_peak, 'unit': 'GB/s', 'frac': kb_gbs / hbm_peak, 'traffic': ncu_traffic('kb_block2_fused') if nb == 8 else None,
                        'peak_kind': peak_kind,
                        'algorithmic_bytes_per_launch': kb_bytes, 'ms_per_launch': kb_ms,
                        'tensor_tflops_bf16_issued': 3.0 * kb_flops / (kb_ms * 1e-3) / 1e12 if kb_ms > 0 else 0.0,
                        'tensor_peak_tflops': tf_sus},
        'tma_engine_launches_total': int(tma_launches),
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        fwd = cpu_forward_fn()
        cb = args.cpu_batch
        inp = synth_inputs(cb, 1)
        fwd(*[t[:1] for t in inp])
        t0 = time.perf_counter()
        fwd(*inp)
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': cb / dt, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d frames, one pass after a 1-frame warm-up, oracle port of the reference '
                                          'PyTorch-CPU path' % cb}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
