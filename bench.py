#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
bench.py -- megapixels/s of dense disparity on B200 for the DeepMatching-for-stereo path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the whole path (descriptors -> correlation -> pyramid ->
backtracking -> sub-pixel -> planes/mosaic) over one synthetic scene:
  N = 1 : BASELINE.json configs[1] -- 1024x1024 pair, ws 15, image_size 64 (+/-64 px),
          stride 60 (ex_deepmatching_rawinput.py:28-30) -> 225 tiles, 904x904 output.
  N > 1 : weak scaling -- the scene grows by 900 rows (15 tile rows) per extra rank, each
          rank solves its strip of tile rows, one NCCL gather of the finished strips to rank 0.
`value`  : output megapixels / s with both scenes already resident in HBM.
`e2e`    : the same through the public API (ImageCutSolver / dm_solve_scene_host) from
           pinned host uint8 scenes to host float64 planes, copies inside the timed region.
One JSON line on stdout (rank 0).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WS, T, STRIDE = 15, 64, 60
MODES = ['elevation', 'elevation2']
SUB_PIX = True
BASE_ROWS, ROWS_PER_RANK, COLS = 1024, 900, 1024
METRIC = 'megapixels/sec dense disparity'


def scene_shape(n_gpus):
    return (BASE_ROWS + ROWS_PER_RANK * (n_gpus - 1), COLS)


def workload_name(n_gpus):
    h, w = scene_shape(n_gpus)
    return '%dx%d synthetic pair, ws=%d, image_size=%d (+/-%d px), stride=%d, modes=%s, sub_pix=%s' % (
        h, w, WS, T, T, STRIDE, '+'.join(MODES), SUB_PIX)


def make_scene(n_gpus):
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    return stereo_pair(scene_shape(n_gpus), seed=1, mode='sine', amp=T // 4)


def load_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tensor=d.get('bf16_tflops_sustained', d['bf16_tflops']), source='measured')
    return dict(hbm=6650.0, tensor=1400.0, source='fallback')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms from the last warm-up steps to the
    end of the end-to-end loop (the timed region of a default run is a few tens of ms)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------- CPU leg
def cpu_sample(img1, img2, n_tiles_total, out_px_total, budget_tiles, threads):
    """The oracle (numpy port of the reference) on a bounded sample of this workload's
    tiles, all host threads.  Returns (MP/s extrapolated linearly in tile count, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dm_oracle as O
    e2 = WS - 1
    len1 = (img1.shape[1] - (T + e2)) // STRIDE
    idx = np.linspace(0, n_tiles_total - 1, budget_tiles).astype(int)

    def one(g):
        y, x = STRIDE * (g // len1), STRIDE * (g % len1)
        return O.solve_tile(img1[y:y + T + e2, x:x + T + e2], img2[y:y + T + e2, x:x + T + e2], WS, MODES, SUB_PIX)

    t = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, idx))
    dt = time.perf_counter() - t
    mp = out_px_total * (len(idx) / float(n_tiles_total)) / 1e6
    return mp / dt, dt, len(idx)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    img1, img2 = make_scene(args.gpus)
    e2 = WS - 1
    len0, len1 = (img1.shape[0] - (T + e2)) // STRIDE, (img1.shape[1] - (T + e2)) // STRIDE
    n_tiles = len0 * len1
    out_px = (STRIDE * (len0 - 1) + T) * (STRIDE * (len1 - 1) + T)
    threads = os.cpu_count() or 1
    sample = min(n_tiles, max(threads, 8))          # one tile per host thread and step (~3 s)
    for _ in range(args.warmup):
        cpu_sample(img1, img2, n_tiles, out_px, sample, threads)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, n = cpu_sample(img1, img2, n_tiles, out_px, sample, threads)
        vals.append(v); secs.append(dt)
    value = float(np.sum([out_px * (sample / float(n_tiles)) / 1e6 for _ in vals]) / np.sum(secs))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'MP/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean(secs)), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': workload_name(args.gpus)},
        'cpu_baseline': {'value': value, 'unit': 'MP/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d of %d tiles per step through oracle.solve_tile (numpy port of the reference, '
                                   'float64 pyramid), extrapolated linearly in tile count' % (sample, n_tiles)},
        'e2e': {'value': value, 'unit': 'MP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------- GPU leg
def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.strips import StripSolver, SharedHostMosaic, input_rows
    from deepmatching_stereo_matching_b200.image_cut_solver import ImageCutSolver, pinned_empty

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, '--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)' % (args.gpus, world)
    torch.cuda.set_device(local_rank)
    if world > 1:
        if not os.environ.get('DM_KEEP_NCCL_DEBUG'):
            os.environ['NCCL_DEBUG'] = 'WARN'
            os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')    # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    img1, img2 = make_scene(world)
    h1 = pinned_empty(img1.shape, np.uint8); h1[...] = img1
    h2 = pinned_empty(img2.shape, np.uint8); h2[...] = img2
    d1 = torch.from_numpy(h1).cuda(non_blocking=True)
    d2 = torch.from_numpy(h2).cuda(non_blocking=True)
    solver = StripSolver(img1.shape, [T, T], [STRIDE, STRIDE], WS, 'cv2.TM_CCOEFF_NORMED', MODES, SUB_PIX, fused=args.fused)
    planes = solver.alloc_planes()
    out_px = solver.out_h * solver.out_w

    def step():
        solver.solve_local(d1, d2, planes)
        return solver.gather(planes)

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        full = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    info = solver.info
    launches_per_step = (info.kernel_launches if info is not None else 0)

    # ---- end to end through the public API: pinned host scenes -> host float64 planes
    lo, hi = solver.tile_rows
    a, b = input_rows(lo, hi, STRIDE, T, WS)
    # N > 1: the mosaic is assembled on the HOST -- one page-locked buffer shared by the ranks,
    # every rank reads its own strip back over its own PCIe link (no GPU-side gather, no 8 strips
    # through rank 0's link)
    mosaic = SharedHostMosaic((solver.n_planes, solver.out_h, solver.out_w), np.float64) if world > 1 else None

    def e2e_step():
        if world == 1:
            s = ImageCutSolver(h1, h2, image_size=[T, T], stride=[STRIDE, STRIDE], window_size=WS,
                               degree_map_mode=MODES, sub_pix=SUB_PIX)
            s.log_flg = False
            s.fused = args.fused
            return s()
        d1[a:b].copy_(torch.from_numpy(h1[a:b]), non_blocking=True)
        d2[a:b].copy_(torch.from_numpy(h2[a:b]), non_blocking=True)
        solver.solve_local(d1, d2, planes)
        mosaic.copy_strip(planes, solver.row_ranges[rank])
        torch.cuda.synchronize()
        dist.barrier()                      # every strip has landed: the mosaic is complete for all ranks
        return mosaic.array

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clk = clocks.stop() if clocks else None
    h2d = 2 * (b - a) * img1.shape[1] * world if world > 1 else 2 * img1.size
    d2h = solver.n_planes * out_px * 8
    if mosaic is not None:
        barrier()
        mosaic.close()

    # ---- per-stage device time (CUDA events on the launching stream) for the roofline
    roofline = None
    if rank == 0:
        peaks = load_peaks()

        def stage_times(prm_, reps=3):
            tctx = _native.Context(timing=True)
            for _ in range(2):
                tctx.solve_device(prm_, d1, d2, planes[:-1], planes[-1])
            acc_, launch_, tinfo_ = {}, {}, None
            for _ in range(reps):
                tinfo_ = tctx.solve_device(prm_, d1, d2, planes[:-1], planes[-1])
                st_ms, launch_ = tctx.stage_ms()
                for k, v in st_ms.items():
                    acc_[k] = acc_.get(k, 0.0) + v / reps
            tctx.close()
            return acc_, launch_, tinfo_

        acc, st_launch, tinfo = stage_times(solver.prm)
        tiles = tinfo.n_tiles
        P = T * T
        kpad = ((WS * WS + 63) // 64) * 64
        fused = bool(tinfo.used_fused)
        # algorithmic work per step (DESIGN.md "Kernels")
        flops = 2.0 * WS * WS * P * P * tiles                                  # unpadded K
        lvl = lambda k: P * P / 16.0 ** k                                       # entries of level k per tile
        first = 1 if fused else 0
        agg_bytes = sum(4.0 * (lvl(k) + lvl(k + 1)) for k in range(first, info.levels - 1)) * tiles
        work = {
            'descriptors': ('hbm', 2.0 * (P * kpad * 2 + (T + WS - 1) ** 2) * tiles),
            'correlation': ('tensor', flops),
            # fused: pooled raw map read + level 1 written; materialising: level 0 read + written
            'normalize': ('hbm', (4.0 * (lvl(0) / 4 + lvl(1)) if fused else 8.0 * lvl(0)) * tiles),
            'aggregate': ('hbm', agg_bytes),
        }
        kern_name = {'descriptors': 'dm_descriptor_row_kernel', 'correlation': 'dm_correlation_umma_kernel',
                     'normalize': 'dm_aggregate_first_kernel' if fused else 'dm_minmax_rectify_kernel',
                     'aggregate': 'dm_aggregate_kernel', 'backtrack': 'dm_backtrack_kernel',
                     'planes': 'dm_final_quad_kernel' if fused else 'dm_planes_kernel'}
        traffic = {}
        tp = os.path.join(REPO, 'profiles', 'ncu_traffic.json')       # dram bytes per launch from ncu --set full
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get('fused' if fused else 'materialising', {})
        kernels = {}
        for k, (bound, w) in work.items():
            if acc.get(k, 0) <= 0:
                continue
            if bound == 'tensor':
                ach, pk, unit = w / (acc[k] * 1e-3) / 1e12, peaks['tensor'], 'TFLOP/s'
            else:
                ach, pk, unit = w / (acc[k] * 1e-3) / 1e9, peaks['hbm'], 'GB/s'
            kernels[k] = {'kernel': kern_name[k], 'bound': bound, 'achieved': ach, 'peak': pk, 'unit': unit, 'frac': ach / pk,
                          'ms': acc[k], 'launches': st_launch[k], 'traffic': traffic.get(kern_name[k])}
            if bound == 'hbm' and ach > pk:
                # the measured peak is a COPY (one read per write); a stream that reads four bytes per byte
                # written pays fewer bus turnarounds and can exceed it (nominal HBM3e: ~7.7 TB/s)
                kernels[k]['note'] = 'read-dominated stream above the measured copy bandwidth (denominator is a 1:1 copy)'
        total = sum(acc.values())
        top = max(kernels, key=lambda k: kernels[k]['ms'])
        roofline = dict(kernels[top])
        roofline.pop('ms'); roofline.pop('launches')
        roofline.update({'peak_source': peaks['source'] + ' (sustained bf16 / copy bandwidth of MEASURED_PEAKS.json)',
                         'share_of_step': acc[top] / total if total > 0 else None,
                         'stage_ms': {k: round(v, 4) for k, v in acc.items()}, 'stage_launches': st_launch,
                         'kernels': kernels, 'used_fused': fused})
        if fused and not args.no_materialising:
            # the stand-alone aggregation kernel on level 0 -> 1 only runs on the materialising path
            import ctypes
            prm0 = type(solver.prm)()
            ctypes.memmove(ctypes.byref(prm0), ctypes.byref(solver.prm), ctypes.sizeof(prm0))
            prm0.fused = 0
            acc0, launch0, _ = stage_times(prm0, reps=2)
            b0 = sum(4.0 * (lvl(k) + lvl(k + 1)) for k in range(0, info.levels - 1)) * tiles
            roofline['materialising_path'] = {
                'stage_ms': {k: round(v, 4) for k, v in acc0.items()},
                'aggregate': {'kernel': 'dm_aggregate_kernel', 'bound': 'hbm', 'achieved': b0 / (acc0['aggregate'] * 1e-3) / 1e9,
                              'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': b0 / (acc0['aggregate'] * 1e-3) / 1e9 / peaks['hbm'],
                              'launches': launch0['aggregate']}}

    if rank == 0:
        value = out_px * args.steps / 1e6 / (ms * 1e-3)
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = min(info.n_tiles, max(32, threads))
            v, dt, n = cpu_sample(img1, img2, info.n_tiles, out_px, sample, threads)
            cpu = {'value': v, 'unit': 'MP/s', 'cores': threads, 'kind': 'port',
                   'sample': '%d of %d tiles through oracle.solve_tile (numpy port of the reference) in %.1f s, '
                             'extrapolated linearly in tile count' % (n, info.n_tiles, dt)}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'MP/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16 operands (exact integers) / f32 accumulate + pyramid / f64 planes', 'data': 'synthetic',
            'config': {'workload': workload_name(world), 'tiles': int(solver.len0 * solver.len1), 'output': [solver.out_h, solver.out_w],
                       'l2': 'no flush: the per-step working set (%.1f GB of pyramid levels) is far larger than the 126 MB L2'
                             % (solver.ctx.workspace_bytes / 1e9),
                       'parallelism': 'tile-row strips x%d, one NCCL gather of the finished strips to rank 0' % world},
            'e2e': {'value': out_px * args.steps / 1e6 / e2e_s, 'unit': 'MP/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'gpu_launches': int(launches_per_step * args.steps),
            'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--fused', type=int, default=-1, help='-1 auto, 0 materialising path, 1 fused tcgen05 path')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-materialising', action='store_true', help='skip the extra stage timing of the materialising path')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
