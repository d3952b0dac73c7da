#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
bench.py -- megapixels/s of dense disparity on B200 for the DeepMatching-for-stereo path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the whole path (descriptors -> correlation -> pyramid -> backtracking ->
sub-pixel -> planes / mosaic) over one synthetic workload of BASELINE.json:

  c2 (default at N = 1)  1024x1024 pair, ws 15, image_size 64 (+/-64 px), stride 60
                         (ex_deepmatching_rawinput.py:28-30) -> 225 tiles, 904x904 output
  c3 (default at N > 1)  4096x4096 scene, same tiling -> 66x66 tiles, 3964x3964 output; STRONG
                         scaling: the 4356 tiles (row-major) are cut into N contiguous ranges, one per rank
  c4                     64 pairs of 512x512, ws 5, image_size 32, stride 32, sub_pix=True, then
                         sub_pix_cal(elevation, score, 1) and (elevation2, score, 0); whole pairs per rank
  c5                     8192x8192 scene written as a headerless .raw file and read back with
                         RawRead.read, ws 15, image_size 128 (+/-128 px), stride 124 -> 64x64 tiles

`value`  : output megapixels / s with the scenes already resident in HBM (CUDA events, max over
           ranks).  N > 1: every rank streams the finished row bands of its strip into ONE mosaic in
           rank 0's memory over NVLink peer memory while it is still solving (`gather` names it;
           `gather_nccl` is the same step with one NCCL gather of the strips after the solve).
`e2e`    : the same through the public API from page-locked host uint8 scenes to host float64
           planes, host<->device copies inside the timed region.  N = 1: ImageCutSolver(...)().
           N > 1: every rank uploads the input rows of its strip and streams its finished rows over
           its own PCIe link into one shared page-locked host mosaic; a step ends when every rank
           has published its strip (a flag per rank in the shared segment).
`parity` : computed outside the timed region -- evenly spaced tiles of the step's own output
           against oracle.solve_tile on the same pixels; the run fails (rc 3) above 1e-3.
One JSON line on stdout (rank 0), printed last.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'megapixels/sec dense disparity'
MODES = ['elevation', 'elevation2']
SUB_PIX = True
FEATURE = 'cv2.TM_CCOEFF_NORMED'

CONFIGS = {
    'c1': dict(shape=(256, 256), ws=5, T=32, stride=32, seed=0, batch=1),       # BASELINE configs[0]: the reference's own CPU-runnable case (a parity case, tests/golden/solver_c1_*)
    'c2': dict(shape=(1024, 1024), ws=15, T=64, stride=60, seed=1, batch=1),
    'c3': dict(shape=(4096, 4096), ws=15, T=64, stride=60, seed=2, batch=1),
    'c4': dict(shape=(512, 512), ws=5, T=32, stride=32, seed=100, batch=64, post=True),
    'c5': dict(shape=(8192, 8192), ws=15, T=128, stride=124, seed=3, batch=1, raw=True),
}
PARITY_TOL = 1e-3


def pick_config(args):
    return args.config or ('c2' if args.gpus == 1 else 'c3')


def geometry(c):
    e2 = c['ws'] - 1
    len0 = (c['shape'][0] - (c['T'] + e2)) // c['stride']
    len1 = (c['shape'][1] - (c['T'] + e2)) // c['stride']
    out = (c['stride'] * (len0 - 1) + c['T'], c['stride'] * (len1 - 1) + c['T'])
    return len0, len1, out


def workload_name(name):
    c = CONFIGS[name]
    h, w = c['shape']
    head = '%dx%d synthetic pair' % (h, w) if c['batch'] == 1 else 'batch of %d synthetic %dx%d pairs' % (c['batch'], h, w)
    tail = ', then sub_pix_cal on both planes' if c.get('post') else ''
    src = ' (headerless .raw file read with RawRead.read)' if c.get('raw') else ''
    return '%s%s, ws=%d, image_size=%d (+/-%d px), stride=%d, modes=%s, sub_pix=%s%s' % (
        head, src, c['ws'], c['T'], c['T'], c['stride'], '+'.join(MODES), SUB_PIX, tail)


def config_dict(name, n_gpus):
    """Identical in both arms (the driver compares them)."""
    c = CONFIGS[name]
    len0, len1, out = geometry(c)
    if c['batch'] > 1:
        par = 'whole pairs per GPU (%d pairs over %d)' % (c['batch'], n_gpus)
    else:
        par = '%d x %d tiles in row-major order cut into %d contiguous ranges, one per GPU' % (len0, len1, n_gpus)
    return {'workload': workload_name(name), 'name': name, 'tiles': int(len0 * len1 * c['batch']), 'output': [int(c['batch']), int(out[0]), int(out[1])],
            'l2': 'no flush: a step streams GBs of pyramid levels through the 126 MB L2 (working set >> L2)',
            'parallelism': par}


def make_scene(name):
    """-> (img1, img2) uint8, shape (S0,S1) or (batch,S0,S1)."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    c = CONFIGS[name]
    if c['batch'] == 1:
        return stereo_pair(c['shape'], seed=c['seed'], mode='sine', amp=c['T'] // 4)
    base = [stereo_pair(c['shape'], seed=c['seed'] + b, mode='sine', amp=c['T'] // 4) for b in range(8)]   # 8 distinct pairs, cycled
    i1 = np.stack([base[b % 8][0] for b in range(c['batch'])])
    i2 = np.stack([base[b % 8][1] for b in range(c['batch'])])
    return i1, i2


def load_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tensor=d.get('bf16_tflops_sustained', d['bf16_tflops']), tensor_burst=d['bf16_tflops'], source='measured (MEASURED_PEAKS.json)')
    # /opt/skills/guides/B200_PROFILING.md fallback figures
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms from the warm-up to the end of the
    measurements (the K timed steps, the sustained block and the end-to-end loop)."""
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


# ------------------------------------------------------------------------------- CPU legs
def tile_inputs(name, img1, img2, g):
    """The two (T+ws-1)^2 tiles of global tile index g (scene-major, then row-major)."""
    c = CONFIGS[name]
    len0, len1, _ = geometry(c)
    e2, T, s = c['ws'] - 1, c['T'], c['stride']
    sc, r = divmod(int(g), len0 * len1)
    gi, gj = divmod(r, len1)
    a, b = (img1[sc], img2[sc]) if img1.ndim == 3 else (img1, img2)
    y, x = s * gi, s * gj
    return a[y:y + T + e2, x:x + T + e2], b[y:y + T + e2, x:x + T + e2], (sc, gi, gj)


def cpu_sample(name, img1, img2, budget_tiles, threads):
    """The oracle (numpy port of the reference) on a bounded, evenly spaced sample of this workload's
    tiles, one tile per host thread.  Returns (MP/s extrapolated linearly in tile count, seconds, n)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dm_oracle as O
    c = CONFIGS[name]
    len0, len1, out = geometry(c)
    n_tiles = len0 * len1 * c['batch']
    out_px = out[0] * out[1] * c['batch']
    idx = np.unique(np.linspace(0, n_tiles - 1, budget_tiles).astype(int))

    def one(g):
        a, b, _ = tile_inputs(name, img1, img2, g)
        d, s = O.solve_tile(a, b, c['ws'], MODES, SUB_PIX)
        if c.get('post'):       # config 4: the post-hoc refinement runs on the mosaic; its cost per tile is the same arithmetic
            O.sub_pix_cal(d[0], s, direction=1)
            O.sub_pix_cal(d[1], s, direction=0)
        return d

    t = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, idx))
    dt = time.perf_counter() - t
    mp = out_px * (len(idx) / float(n_tiles)) / 1e6
    return mp / dt, dt, len(idx)


def cpu_budget(name, threads):
    """(tiles per sample, threads): a T=128 tile holds a 2 GB float64 level 0 -- few at a time."""
    if CONFIGS[name]['T'] >= 128:
        return 1, 1
    return max(threads, 8), threads


def reference_unmodified(name, img1, img2, n_sample=8):
    """The UNMODIFIED reference (ImageCutSolver._solver on a sample of tiles) when its checkout is
    importable on this machine (build container: /root/reference; GPU box: baseline/_ref if the driver
    put one there).  Returns a cpu_baseline-style dict or None."""
    for root in (os.path.join(REPO, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(root, 'misc')):
            break
    else:
        return None
    code = r'''
import sys, time, json, io, contextlib
sys.path.insert(0, %r)
import numpy as np
from misc.image_cut_solver import ImageCutSolver
d = np.load(sys.argv[1])
c = json.loads(sys.argv[2])
t0 = time.perf_counter()
n = 0
with contextlib.redirect_stdout(io.StringIO()):
    for k in range(d['a'].shape[0]):
        s = ImageCutSolver(d['a'][k], d['b'][k], image_size=[c['T'], c['T']], stride=[c['stride'], c['stride']], window_size=c['ws'],
                           degree_map_mode=c['modes'], sub_pix=True)
        s.log_flg = False
        s._solver(d['a'][k], d['b'][k])
        n += 1
print(json.dumps({'tiles': n, 'seconds': time.perf_counter() - t0}))
''' % root
    c = CONFIGS[name]
    len0, len1, out = geometry(c)
    n_tiles = len0 * len1 * c['batch']
    if c['T'] >= 128:
        n_sample = 1
    idx = np.unique(np.linspace(0, n_tiles - 1, n_sample).astype(int))
    tiles = [tile_inputs(name, img1, img2, g) for g in idx]
    with tempfile.TemporaryDirectory() as td:
        np.savez(os.path.join(td, 't.npz'), a=np.stack([t[0] for t in tiles]), b=np.stack([t[1] for t in tiles]))
        try:
            r = subprocess.run([sys.executable, '-c', code, os.path.join(td, 't.npz'),
                                json.dumps({'T': c['T'], 'stride': c['stride'], 'ws': c['ws'], 'modes': MODES})],
                               capture_output=True, text=True, timeout=900, cwd=td)
            res = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as exc:
            return {'unavailable': 'unmodified reference at %s did not run: %s' % (root, exc)}
    mp = out[0] * out[1] * c['batch'] * (res['tiles'] / float(n_tiles)) / 1e6
    return {'value': mp / res['seconds'], 'unit': 'MP/s', 'cores': os.cpu_count() or 1, 'kind': 'reference',
            'sample': '%d of %d tiles through the unmodified reference ImageCutSolver._solver (%s; its joblib n_jobs=-1 and OpenCV threads), '
                      'one tile after the other in %.1f s, extrapolated linearly in tile count' % (res['tiles'], n_tiles, root, res['seconds'])}


def run_reference(args):
    """The reference's CPU implementation of the path on the box's host cores.  The reference is
    Python + OpenCV + torch and is not shipped to the GPU box, so the timed arm is the numpy port
    (oracle/, kind "port", ~20x FASTER than the unmodified reference); where
    the reference itself is importable its unmodified _solver is timed beside it."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    name = pick_config(args)
    c = CONFIGS[name]
    img1, img2 = make_scene(name)
    threads = os.cpu_count() or 1
    sample, workers = cpu_budget(name, threads)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(name, img1, img2, sample, workers)
    mps, secs = [], []
    for _ in range(args.steps):
        v, dt, n = cpu_sample(name, img1, img2, sample, workers)
        mps.append(v * dt); secs.append(dt)
    value = float(np.sum(mps) / np.sum(secs))
    len0, len1, out = geometry(c)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'MP/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean(secs)), 'higher_is_better': True,
        'scaling': 'weak' if args.gpus == 1 else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(name, args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'MP/s', 'cores': workers, 'kind': 'port',
                         'sample': '%d of %d tiles per step through oracle.solve_tile (numpy port of the reference, float64 pyramid; '
                                   '~20x faster than the unmodified reference), extrapolated linearly in tile count' % (n, len0 * len1 * c['batch'])},
        'e2e': {'value': value, 'unit': 'MP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if not args.no_unmodified:
        ref = reference_unmodified(name, img1, img2)
        if ref is not None:
            line['reference_unmodified'] = ref
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------- parity
def parity_block(name, img1, img2, planes, n_tiles_check, n_scenes=None):
    """Evenly spaced tiles of the step's own output against oracle.solve_tile on the same pixels
    (only the pixels a tile owns: later tiles overwrite the overlap, misc/image_cut_solver.py:165-175).
    planes: (batch, n_modes + 1, out_h, out_w) float64 host array."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dm_oracle as O
    c = CONFIGS[name]
    len0, len1, _ = geometry(c)
    T, s = c['T'], c['stride']
    n_tiles = len0 * len1 * (c['batch'] if n_scenes is None else n_scenes)     # the pairs rank 0 holds
    idx = np.unique(np.linspace(0, n_tiles - 1, n_tiles_check).astype(int))
    nm = len(MODES)

    def one(g):
        a, b, (sc, gi, gj) = tile_inputs(name, img1, img2, g)
        rd, rs = O.solve_tile(a, b, c['ws'], MODES, SUB_PIX)
        o0 = T if gi == len0 - 1 else min(T, s)
        o1 = T if gj == len1 - 1 else min(T, s)
        got = planes[sc, :, s * gi:s * gi + o0, s * gj:s * gj + o1]
        d, sco = got[:nm], got[nm]
        rd, rs = rd[:, :o0, :o1], rs[:o0, :o1]
        ok = ~(np.isnan(d) | np.isnan(rd))
        int_bad = (np.abs(d - rd) > 0.5) & ok
        rel = np.abs(d - rd) / np.maximum(1.0, np.abs(rd))
        same = ok & ~int_bad
        sok = ~(np.isnan(sco) | np.isnan(rs)) & ~int_bad.any(0)
        sdiff = np.abs(sco - rs)
        return dict(n=int(ok.sum()), int_bad=int(int_bad.sum()), sub_bad=int(((rel > PARITY_TOL) & same).sum()),
                    max_rel=float(rel[same].max()) if same.any() else 0.0, n_score=int(sok.sum()),
                    score_bad=int((sdiff[sok] > PARITY_TOL).sum()), score_max=float(sdiff[sok].max()) if sok.any() else 0.0,
                    nan_mismatch=int((np.isnan(d) != np.isnan(rd)).sum()))

    workers = 1 if T >= 128 else min(len(idx), os.cpu_count() or 1)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as ex:
        res = list(ex.map(one, idx))
    n = sum(r['n'] for r in res)
    ns = sum(r['n_score'] for r in res)
    out = {
        'tiles_checked': int(len(idx)), 'tiles_total': int(n_tiles), 'pixels_checked': int(n),
        'int_disagreement': sum(r['int_bad'] for r in res) / max(1.0, float(n)),
        'subpix_frac_above_tol': sum(r['sub_bad'] for r in res) / max(1.0, float(n)),
        'max_subpix_rel': max(r['max_rel'] for r in res),
        'score_frac_above_tol': sum(r['score_bad'] for r in res) / max(1.0, float(ns)),
        'score_max_abs': max(r['score_max'] for r in res),
        'nan_mismatch': sum(r['nan_mismatch'] for r in res),
        'tol': PARITY_TOL, 'against': 'oracle.solve_tile (numpy restatement of the reference, float64 pyramid)',
        'seconds': round(time.perf_counter() - t0, 1),
    }
    out['ok'] = bool(out['int_disagreement'] <= PARITY_TOL and out['subpix_frac_above_tol'] <= PARITY_TOL
                     and out['score_frac_above_tol'] <= PARITY_TOL and out['nan_mismatch'] == 0)
    return out


# ------------------------------------------------------------------------------- GPU leg
def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.strips import StripSolver, SharedHostMosaic, PeerMosaic, input_rows
    from deepmatching_stereo_matching_b200.image_cut_solver import ImageCutSolver, pinned_empty, solve_batch
    from deepmatching_stereo_matching_b200.raw_read import RawRead

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, '--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)' % (args.gpus, world)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    name = pick_config(args)
    c = CONFIGS[name]
    T, WS, STRIDE, B = c['T'], c['ws'], c['stride'], c['batch']
    img1, img2 = make_scene(name)
    raw_paths = None
    if c.get('raw'):
        # the scene exists as a headerless .raw file, as ex_deepmatching_rawinput.py:50-55 expects it
        td = tempfile.mkdtemp(prefix='dm_raw_')
        raw_paths = (os.path.join(td, 'band1_r%d.raw' % rank), os.path.join(td, 'band2_r%d.raw' % rank))
        img1.tofile(raw_paths[0]); img2.tofile(raw_paths[1])
        img1 = RawRead.read(raw_paths[0], size=(c['shape'][1], c['shape'][0]))
        img2 = RawRead.read(raw_paths[1], size=(c['shape'][1], c['shape'][0]))
    h1 = pinned_empty(img1.shape, np.uint8); h1[...] = img1
    h2 = pinned_empty(img2.shape, np.uint8); h2[...] = img2
    len0, len1, out_hw = geometry(c)
    out_px = out_hw[0] * out_hw[1] * B
    nm = len(MODES)
    lib = _native.lib()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    peer = None
    checks = {}

    if B == 1:
        # ---------------------------------------------------------------- one scene: strips of tile rows
        solver = StripSolver(c['shape'], [T, T], [STRIDE, STRIDE], WS, FEATURE, MODES, SUB_PIX, fused=args.fused)
        a, b = solver.input_rows
        d1 = torch.zeros(c['shape'], dtype=torch.uint8, device='cuda')
        d2 = torch.zeros(c['shape'], dtype=torch.uint8, device='cuda')
        d1[a:b].copy_(torch.from_numpy(h1[a:b])); d2[a:b].copy_(torch.from_numpy(h2[a:b]))      # a rank holds only the rows its strip reads
        planes = solver.alloc_planes()
        peer, peer_error = None, None
        if world > 1:
            # the mosaic in rank 0's memory, mapped into every rank through CUDA IPC; if the box does not allow that
            # (no peer access, IPC disabled in a container) every rank falls back to the NCCL gather for `value`
            try:
                if os.environ.get('DM_BENCH_NO_PEER'):
                    raise RuntimeError('DM_BENCH_NO_PEER is set')
                peer = PeerMosaic((solver.n_planes, solver.out_h, solver.out_w))
            except Exception as exc:                 # noqa: BLE001 -- reported in the line, not swallowed
                peer_error = '%s: %s' % (type(exc).__name__, exc)
            flags = [None] * world
            dist.all_gather_object(flags, peer_error)
            if any(f is not None for f in flags):
                if peer is not None:
                    peer.close()
                peer = None
                peer_error = next(f for f in flags if f is not None)
        gather_name = 'none (one GPU)' if world == 1 else (
            'finished row bands streamed into rank 0\'s mosaic over NVLink peer memory (CUDA IPC) while the strip is still being solved'
            if peer is not None else 'NCCL point-to-point gather of the owned rows after the solve (peer mosaic unavailable: %s)' % peer_error)

        def step():
            if world == 1:
                solver.solve_local(d1, d2, planes)
                return planes
            if peer is None:
                return step_nccl()
            solver.solve_into(d1, d2, planes, peer)
            return peer.tensor

        def step_nccl():
            solver.solve_local(d1, d2, planes)
            return solver.gather(planes)

        def result_planes(t):
            return t.cpu().numpy()[None] if t is not None else None
        info_of = lambda: solver.info
        solve_prm = solver.prm
    else:
        # ---------------------------------------------------------------- batch of pairs: whole pairs per rank
        parts = _native.partition_tile_rows(B, world)
        p_lo, p_hi = parts[rank]
        nb = p_hi - p_lo
        d1 = torch.from_numpy(h1[p_lo:p_hi]).cuda(); d2 = torch.from_numpy(h2[p_lo:p_hi]).cuda()
        prm = _native.scene_params(c['shape'], [T, T], [STRIDE, STRIDE], WS, FEATURE, MODES, SUB_PIX, None, args.fused, n_scenes=nb)
        ctx = _native.Context()
        dmap = torch.zeros((nb, nm, out_hw[0], out_hw[1]), dtype=torch.float64, device='cuda')
        omap = torch.zeros((nb, out_hw[0], out_hw[1]), dtype=torch.float64, device='cuda')
        post = torch.empty_like(dmap)
        holder = {}
        gather_name = 'none (whole pairs per GPU, results stay on their GPU)'

        def post_refine():      # config 4: sub_pix_cal(elevation, score, direction=1), (elevation2, score, direction=0)  (image_cut_solver.py:137)
            for bb in range(nb):
                for m, mode in enumerate(MODES):
                    _native.check(lib.dm_sub_pix_cal(_native.ptr(dmap[bb, m]), _native.ptr(omap[bb]), out_hw[0], out_hw[1],
                                                     1 if mode == 'elevation' else 0, 100.0, _native.ptr(post[bb, m]), _native.stream_ptr()))

        def step():
            holder['info'] = ctx.solve_device(prm, d1, d2, dmap, omap)
            if c.get('post'):
                post_refine()
            return dmap

        step_nccl = None

        def result_planes(t):
            return torch.cat([dmap, omap[:, None]], 1).cpu().numpy()
        info_of = lambda: holder.get('info')
        solve_prm = prm

    # ---------------------------------------------------------------- device-resident: K timed steps
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        full = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    info = info_of()
    local_launches = info.kernel_launches if info is not None else 0
    if c.get('post'):
        local_launches += 2 * (d1.shape[0])
    launches_per_step = float(local_launches)
    if world > 1:
        t = torch.tensor([float(local_launches)], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches_per_step = float(t.item())
    got_planes = result_planes(full) if rank == 0 else None       # the step's own output, for the parity block

    # ---- sustained: the same step back to back for >= args.sustain seconds (clock record, power-capped rate)
    sustained = None
    if args.sustain > 0:
        n_s = max(args.steps, int(np.ceil(args.sustain * 1e3 / max(ms / args.steps, 1e-3))))
        n_s = int(max_over_ranks(n_s))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_s):
            step()
        s1.record()
        barrier()
        sms = max_over_ranks(s0.elapsed_time(s1))
        sustained = {'value': out_px * n_s / 1e6 / (sms * 1e-3), 'unit': 'MP/s', 'steps': n_s, 'seconds': sms * 1e-3, 'ms_per_step': sms / n_s}

    # ---- the same step with one NCCL gather of the strips after the solve
    gather_nccl = None
    if world > 1 and step_nccl is not None and not args.no_nccl_gather:
        for _ in range(2):
            step_nccl()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            step_nccl()
        g1.record()
        barrier()
        gms = max_over_ranks(g0.elapsed_time(g1))
        checks['nccl_gather_equals_streamed_mosaic'] = bool(torch.equal(planes, peer.tensor)) if (rank == 0 and peer is not None) else None
        gather_nccl = {'value': out_px * args.steps / 1e6 / (gms * 1e-3), 'unit': 'MP/s', 'ms_per_step': gms / args.steps,
                       'how': 'solve the strip, then point-to-point NCCL transfers of the owned rows (float64) to rank 0'}

    # ---- strong scaling denominator: the whole workload on rank 0 alone, in the same run
    single = None
    if world > 1 and B == 1 and not args.no_single:
        if rank == 0:
            dd1 = torch.from_numpy(h1).cuda(); dd2 = torch.from_numpy(h2).cuda()
            sprm = _native.scene_params(c['shape'], [T, T], [STRIDE, STRIDE], WS, FEATURE, MODES, SUB_PIX, None, args.fused)
            sctx = _native.Context()
            for _ in range(2):
                sctx.solve_device(sprm, dd1, dd2, planes[:-1], planes[-1])
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(3, min(args.steps, 10))
            q0.record()
            for _ in range(reps):
                sctx.solve_device(sprm, dd1, dd2, planes[:-1], planes[-1])
            q1.record()
            torch.cuda.synchronize()
            qms = q0.elapsed_time(q1) / reps
            if peer is not None:
                checks['strips_equal_single_gpu_solve'] = bool(torch.equal(planes, peer.tensor))      # bit for bit
            elif got_planes is not None:
                checks['strips_equal_single_gpu_solve'] = bool(np.array_equal(planes.cpu().numpy()[None], got_planes, equal_nan=True))
            single = {'value': out_px / 1e6 / (qms * 1e-3), 'unit': 'MP/s', 'ms_per_step': qms, 'n_gpus': 1,
                      'how': 'the same workload solved by rank 0 alone (device-resident), for the strong-scaling ratio'}
            sctx.close()
            del dd1, dd2
        barrier()

    # ---------------------------------------------------------------- end to end through the public API
    mosaic = None
    if B == 1 and world > 1:
        mosaic = SharedHostMosaic((solver.n_planes, solver.out_h, solver.out_w), np.float64)
    e2e_counter = [0]
    e2e_stamps = []          # N > 1: per step (this rank's own upload + solve + copies, then waiting for the slowest rank)

    def e2e_step():
        if B == 1 and world == 1:
            s = ImageCutSolver(h1, h2, image_size=[T, T], stride=[STRIDE, STRIDE], window_size=WS, degree_map_mode=MODES, sub_pix=SUB_PIX)
            s.log_flg = False
            s.fused = args.fused
            s.devices = [local_rank]                 # --gpus 1 means one GPU, whatever the box shows
            return s()
        if B == 1:
            e2e_counter[0] += 1
            ta_ = time.perf_counter()
            solver.solve_host_into(h1, h2, mosaic)   # upload of the strip's rows, solve, finished rows streamed into the shared host mosaic
            tb_ = time.perf_counter()
            mosaic.publish(e2e_counter[0])
            mosaic.wait(e2e_counter[0])              # the whole mosaic has landed
            e2e_stamps.append((tb_ - ta_, time.perf_counter() - tb_))
            return mosaic.array
        dm_, om_ = solve_batch(h1[p_lo:p_hi], h2[p_lo:p_hi], image_size=[T, T], stride=[STRIDE, STRIDE], window_size=WS,
                               degree_map_mode=MODES, sub_pix=SUB_PIX, fused=args.fused, devices=[local_rank])
        if c.get('post'):
            from deepmatching_stereo_matching_b200.sub_pix_cal import sub_pix_cal_batch
            return sub_pix_cal_batch(dm_, om_, [1 if m == 'elevation' else 0 for m in MODES]), om_
        return dm_, om_

    e2e_steps = args.steps if (ms / args.steps) < 50 else max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {'value': out_px * e2e_steps / 1e6 / e2e_s, 'unit': 'MP/s', 'steps': e2e_steps, 'ms_per_step': 1e3 * e2e_s / e2e_steps}
    if world > 1 and B == 1:
        # where an end-to-end step goes: every rank's own part (upload + solve + streamed copies) and its wait for the
        # slowest rank, over the timed steps -- event stamps instead of a profiler timeline
        own = float(np.mean([x[0] for x in e2e_stamps[-e2e_steps:]])) * 1e3
        wait = float(np.mean([x[1] for x in e2e_stamps[-e2e_steps:]])) * 1e3
        allv = [None] * world
        dist.all_gather_object(allv, (own, wait, int(solver.tiles[1] - solver.tiles[0])))
        if rank == 0:
            e2e['timeline_ms'] = {'own_part_per_rank': [round(v[0], 3) for v in allv], 'wait_for_slowest_per_rank': [round(v[1], 3) for v in allv],
                                  'tiles_per_rank': [v[2] for v in allv],
                                  'reading': 'a step lasts max(own part) + the flag round; ranks that finish early wait'}
    if B == 1:
        rows_in = [input_rows(l // len1, (h - 1) // len1 + 1, STRIDE, T, WS) for (l, h) in solver.tile_parts if h > l]
        e2e['h2d_bytes_per_step'] = int(sum(2 * (bb - aa) * c['shape'][1] for aa, bb in rows_in))
        e2e['d2h_bytes_per_step'] = int((nm + 1) * out_px * 8)
    else:
        e2e['h2d_bytes_per_step'] = int(2 * h1.size)
        e2e['d2h_bytes_per_step'] = int((nm + 1) * out_px * 8 + (nm * out_px * 8 if c.get('post') else 0))
        if c.get('post'):
            e2e['h2d_bytes_per_step'] += int((nm + 1) * out_px * 8)     # sub_pix_cal takes host arrays (misc/sub_pix_cal.py:22)
    if raw_paths is not None and world == 1:
        # C5 from the .raw files: RawRead.read (host, misc/raw_read.py:36-45) + the solve
        t0 = time.perf_counter()
        a_ = RawRead.read(raw_paths[0], size=(c['shape'][1], c['shape'][0]))
        b_ = RawRead.read(raw_paths[1], size=(c['shape'][1], c['shape'][0]))
        read_s = time.perf_counter() - t0
        s = ImageCutSolver(a_, b_, image_size=[T, T], stride=[STRIDE, STRIDE], window_size=WS, degree_map_mode=MODES, sub_pix=SUB_PIX)
        s.log_flg = False; s.devices = [local_rank]
        s()
        tot = time.perf_counter() - t0
        e2e['from_raw_file'] = {'value': out_px / 1e6 / tot, 'unit': 'MP/s', 'raw_read_s': read_s, 'total_s': tot,
                                'how': 'RawRead.read of both 64 MB files (pageable numpy arrays) + ImageCutSolver(...)()'}
    clk = clocks.stop() if clocks else None
    if mosaic is not None:
        barrier()
        if rank == 0:
            ref_ = peer.tensor.cpu().numpy() if peer is not None else got_planes[0]
            checks['host_mosaic_equals_streamed_mosaic'] = bool(np.array_equal(mosaic.array, ref_, equal_nan=True))
        mosaic.close()
    if raw_paths is not None:
        for p_ in raw_paths:
            try:
                os.remove(p_)
            except OSError:
                pass

    # ---------------------------------------------------------------- per-stage device time -> roofline
    roofline = None
    if rank == 0:
        peaks = load_peaks()
        if B == 1:
            t_img1, t_img2, t_dm, t_om = d1, d2, planes[:-1], planes[-1]
        else:
            t_img1, t_img2, t_dm, t_om = d1, d2, dmap, omap

        def stage_times(prm_, seconds):
            """stage events inside a back-to-back run of at least `seconds` (the power-capped rate, the one the
            sustained peak is the denominator for); averaged over the second half of the run"""
            tctx = _native.Context(timing=True)
            for _ in range(2):
                tinfo_ = tctx.solve_device(prm_, t_img1, t_img2, t_dm, t_om)
            torch.cuda.synchronize()
            one = max(ms / args.steps, 0.05)
            reps = int(max(4, min(400, np.ceil(seconds * 1e3 / one))))
            acc_, launch_, cnt = {}, {}, 0
            for r_ in range(reps):
                tinfo_ = tctx.solve_device(prm_, t_img1, t_img2, t_dm, t_om)
                if r_ >= reps // 2:
                    st_ms, launch_ = tctx.stage_ms()
                    cnt += 1
                    for k, v in st_ms.items():
                        acc_[k] = acc_.get(k, 0.0) + v
            tctx.close()
            return {k: v / cnt for k, v in acc_.items()}, launch_, tinfo_

        acc, st_launch, tinfo = stage_times(solve_prm, args.stage_seconds)
        tiles = tinfo.n_tiles
        P = T * T
        kpad = int(lib.dm_kpad(WS))
        fused = bool(tinfo.used_fused)
        L = int(tinfo.levels)
        # algorithmic work per step (DESIGN.md "Kernels")
        flops = 2.0 * WS * WS * P * P * tiles                                  # unpadded K
        lvl = lambda k: P * P / 16.0 ** k                                       # entries of level k per tile
        first = 1 if fused else 0
        agg_bytes = sum(4.0 * (lvl(k) + lvl(k + 1)) for k in range(first, L - 1)) * tiles
        desc_bytes = 2.0 * (P * kpad * 2 + (T + WS - 1) ** 2) * tiles
        # correlation: the descriptors are read once, the pooled raw map (fused) or level 0 is written once
        corr_bytes = (desc_bytes / 2.0 + 4.0 * (lvl(0) / 4 if fused else lvl(0)) * tiles)
        work = {
            'descriptors': [('hbm', desc_bytes)],
            'correlation': [('tensor', flops), ('hbm', corr_bytes)],
            # fused: pooled raw map read + level 1 written; materialising: level 0 read + written
            'normalize': [('hbm', (4.0 * (lvl(0) / 4 + lvl(1)) if fused else 8.0 * lvl(0)) * tiles)],
            'aggregate': [('hbm', agg_bytes)],
        }
        kern_name = {'descriptors': 'dm_descriptor_row_kernel', 'correlation': 'dm_correlation_umma_kernel',
                     'normalize': 'dm_aggregate_first_kernel' if fused else 'dm_minmax_rectify_kernel',
                     'aggregate': 'dm_aggregate_kernel', 'backtrack': 'dm_upper_tail_kernel' if fused else 'dm_backtrack_kernel',
                     'planes': 'dm_final_patch_kernel' if fused else 'dm_planes_kernel'}
        traffic = {}
        tp = os.path.join(REPO, 'profiles', 'ncu_traffic.json')       # dram bytes per launch from ncu --set full (C2 only)
        if os.path.exists(tp) and name == 'c2':
            traffic = json.load(open(tp)).get('fused' if fused else 'materialising', {})
        elif os.path.exists(tp) and fused:
            # other shapes: the correlation kernel's measured DRAM bytes per tile, scaled to the tiles of one launch
            try:
                per_tile = json.load(open(tp)).get('correlation_per_tile', {}).get('t%d_ws%d' % (T, WS))
                n_launch = int(st_launch.get('correlation', 0))
                if per_tile and n_launch > 0:
                    traffic = {'dm_correlation_umma_kernel': float(per_tile) * tiles / n_launch}
            except Exception:
                traffic = {}
        kernels = {}
        for k, cands in work.items():
            if acc.get(k, 0) <= 0:
                continue
            best = None
            for bound, w in cands:
                if bound == 'tensor':
                    ach, pk, unit = w / (acc[k] * 1e-3) / 1e12, peaks['tensor'], 'TFLOP/s'
                else:
                    ach, pk, unit = w / (acc[k] * 1e-3) / 1e9, peaks['hbm'], 'GB/s'
                ent = {'kernel': kern_name[k], 'bound': bound, 'achieved': ach, 'peak': pk, 'unit': unit, 'frac': ach / pk,
                       'ms': acc[k], 'launches': st_launch[k], 'traffic': traffic.get(kern_name[k])}
                if bound == 'tensor':
                    ent['frac_of_burst_peak'] = ach / peaks['tensor_burst']
                if best is None or ent['frac'] > best['frac']:
                    best = ent          # the roof the kernel is closest to is the one that bounds it
            kernels[k] = best
            if best['bound'] == 'hbm' and best['achieved'] > best['peak']:
                # the measured peak is a COPY (one read per write); a stream that reads four bytes per byte
                # written pays fewer bus turnarounds and can exceed it (nominal HBM3e: ~7.7 TB/s)
                best['note'] = 'read-dominated stream above the measured copy bandwidth (denominator is a 1:1 copy)'
        for k in ('backtrack', 'planes'):
            if acc.get(k, 0) > 0:
                kernels[k] = {'kernel': kern_name[k], 'bound': 'latency / issue', 'ms': acc[k], 'launches': st_launch[k]}
        total = sum(acc.values())
        top = max(work, key=lambda k: acc.get(k, 0))
        roofline = dict(kernels[top])
        roofline.pop('ms'); roofline.pop('launches')
        roofline.update({'peak_source': peaks['source'] + ': sustained bf16 / copy bandwidth; stage times are CUDA events on the launching stream '
                                        'inside a %.1f s back-to-back run' % args.stage_seconds,
                         'share_of_step': acc[top] / total if total > 0 else None,
                         'stage_ms': {k: round(v, 4) for k, v in acc.items()}, 'stage_launches': st_launch,
                         'kernels': kernels, 'used_fused': fused})

    # ---------------------------------------------------------------- parity + cpu baseline (outside every timed region)
    parity, cpu = None, None
    if rank == 0:
        if not args.no_parity:
            ntc = args.parity_tiles if args.parity_tiles > 0 else (2 if T >= 128 else 16)
            # config 4: parity is checked on the solver output; the post-hoc refinement is float64 arithmetic
            # that is bit-identical to numpy (tests/test_gpu_parity.py::test_sub_pix_cal_bit_exact)
            parity = parity_block(name, img1, img2, got_planes, ntc, n_scenes=got_planes.shape[0])
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            sample, workers = cpu_budget(name, threads)
            sample = max(sample, 2 * workers) if T < 128 else sample
            v, dt, n = cpu_sample(name, img1, img2, sample, workers)
            cpu = {'value': v, 'unit': 'MP/s', 'cores': workers, 'kind': 'port',
                   'sample': '%d of %d tiles through oracle.solve_tile (numpy port of the reference) in %.1f s, '
                             'extrapolated linearly in tile count' % (n, len0 * len1 * B, dt)}
    if peer is not None:
        barrier()
        peer.close()
    if world > 1:
        # the JSON line must be the LAST line on stdout: with NCCL_DEBUG=INFO every rank still prints while
        # it tears its communicator down, so rank 0 prints only after the other ranks have left
        done_dir = [tempfile.mkdtemp(prefix='dm_done_') if rank == 0 else None]
        dist.broadcast_object_list(done_dir, src=0)
        dist.destroy_process_group()
        sys.stdout.flush(); sys.stderr.flush()
        if rank != 0:
            open(os.path.join(done_dir[0], 'rank%d' % rank), 'w').close()
            os._exit(0)                     # no library destructors: nothing more is printed by this rank
        t_wait = time.perf_counter()
        while len(os.listdir(done_dir[0])) < world - 1 and time.perf_counter() - t_wait < 20:
            time.sleep(0.05)
        time.sleep(0.3)
        import shutil
        shutil.rmtree(done_dir[0], ignore_errors=True)
    rc = 0
    if rank == 0:
        value = out_px * args.steps / 1e6 / (ms * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'MP/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak' if world == 1 else 'strong', 'vs_baseline': None,
            'dtype': 'bf16 operands (exact integers) / f32 accumulate + pyramid / f64 planes', 'data': 'synthetic',
            'config': config_dict(name, world), 'gather': gather_name,
            'e2e': e2e, 'gpu_launches': int(launches_per_step * args.steps),
            'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu, 'parity': parity,
            'sustained': sustained, 'gather_nccl': gather_nccl, 'single_gpu_same_workload': single, 'checks': checks,
        }
        sys.stdout.flush()
        print(json.dumps(line))
        sys.stdout.flush()
        if parity is not None and not parity['ok']:
            sys.stderr.write('PARITY FAILED: %s\n' % json.dumps(parity))
            rc = 3
        if any(v is False for v in checks.values()):
            sys.stderr.write('MULTI-GPU CHECK FAILED: %s\n' % json.dumps(checks))
            rc = 4
        if world > 1:
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(rc)
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default=None, choices=sorted(CONFIGS), help='default: c2 at --gpus 1, c3 (strong scaling) above')
    ap.add_argument('--fused', type=int, default=-1, help='-1 auto, 0 materialising path, 1 fused tcgen05 path')
    ap.add_argument('--sustain', type=float, default=1.0, help='seconds of back-to-back steps for the `sustained` block (0 = skip)')
    ap.add_argument('--stage-seconds', type=float, default=0.5, help='length of the back-to-back run the stage events are taken in')
    ap.add_argument('--parity-tiles', type=int, default=0, help='tiles checked against the oracle (default 16; 2 at image_size 128)')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-nccl-gather', action='store_true', help='skip the extra timing of the NCCL gather variant')
    ap.add_argument('--no-single', action='store_true', help='skip the single-GPU run of the same workload at N > 1')
    ap.add_argument('--no-unmodified', action='store_true', help='reference arm: do not also time the unmodified reference')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
