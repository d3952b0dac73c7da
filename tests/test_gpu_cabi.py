"""A plain C program drives the GPU through include/dmstereo.h: dm_ctx_create,
dm_solve_scene_host (and dm_multi_solve_scene_host) on a synthetic shifted pair."""
import os
import shutil
import subprocess

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dmstereo.h"

/* smooth-ish pseudo-random texture: a sum of a few box-blurred LCG noise fields */
static void texture(unsigned char* t, int h, int w) {
    unsigned int s = 12345u;
    float* f = (float*)malloc(sizeof(float) * h * w);
    for (int i = 0; i < h * w; ++i) { s = s * 1664525u + 1013904223u; f[i] = (float)(s >> 24); }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float a = 0.f; int n = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                    a += f[yy * w + xx]; ++n;
                }
            t[y * w + x] = (unsigned char)(a / n);
        }
    free(f);
}

int main(int argc, char** argv) {
    const int S = 400, SHIFT = 3, multi = argc > 1;
    unsigned char* tex = (unsigned char*)malloc(S * (S + 2 * SHIFT));
    unsigned char* img1 = (unsigned char*)malloc(S * S);
    unsigned char* img2 = (unsigned char*)malloc(S * S);
    texture(tex, S, S + 2 * SHIFT);
    for (int y = 0; y < S; ++y)
        for (int x = 0; x < S; ++x) {
            img2[y * S + x] = tex[y * (S + 2 * SHIFT) + SHIFT + x];
            img1[y * S + x] = tex[y * (S + 2 * SHIFT) + x];          /* img1(x) = img2(x - SHIFT) */
        }
    dm_scene_params p;
    dm_scene_info info;
    memset(&p, 0, sizeof p);
    p.scene_h = S; p.scene_w = S; p.t0 = 32; p.t1 = 32; p.s0 = 30; p.s1 = 30; p.ws = 7;
    p.method = DM_TM_CCOEFF_NORMED; p.n_modes = 2; p.modes[0] = DM_MODE_ELEVATION; p.modes[1] = DM_MODE_ELEVATION2;
    p.sub_pix = 0; p.fused = -1;
    if (dm_scene_geometry(&p, &info) != DM_OK) { printf("geometry: %s\n", dm_last_error()); return 1; }
    const size_t plane = (size_t)info.out_h * info.out_w;
    double* d_map = (double*)malloc(sizeof(double) * plane * 2);
    double* out_map = (double*)malloc(sizeof(double) * plane);
    int rc;
    if (!multi) {
        dm_ctx* ctx = NULL;
        if ((rc = dm_ctx_create(&ctx)) != DM_OK) { printf("ctx: %s\n", dm_last_error()); return 2; }
        rc = dm_solve_scene_host(ctx, &p, img1, img2, d_map, out_map, &info);
        if (rc != DM_OK) { printf("solve: %s\n", dm_last_error()); return 3; }
        dm_ctx_destroy(ctx);
    } else {
        dm_multi* m = NULL;
        if ((rc = dm_multi_create(NULL, 0, &m)) != DM_OK) { printf("multi: %s\n", dm_last_error()); return 2; }
        rc = dm_multi_solve_scene_host(m, &p, 0, img1, img2, d_map, out_map, &info);
        if (rc != DM_OK) { printf("multi solve: %s\n", dm_last_error()); return 3; }
        printf("devices %d ", dm_multi_device_count(m));
        dm_multi_destroy(m);
    }
    size_t hit = 0;
    for (size_t i = 0; i < plane; ++i) hit += (d_map[i] == (double)SHIFT && d_map[plane + i] == 0.0);
    printf("%d %d %d %d %.4f\n", info.out_h, info.out_w, info.n_tiles, info.kernel_launches, (double)hit / (double)plane);
    return 0;
}
'''


@pytest.mark.parametrize('multi', [False, True])
def test_c_program_solves_a_scene_on_the_gpu(tmp_path, multi):
    import torch
    assert torch.cuda.is_available()
    from deepmatching_stereo_matching_b200 import _native
    _native.lib()
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('gcc not available')
    src = tmp_path / 'solve.c'
    src.write_text(SRC)
    exe = tmp_path / 'solve'
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run([gcc, '-std=c99', '-O1', '-Wall', '-Werror', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe),
                    '-L', libdir, '-ldmstereo', '-Wl,-rpath,' + libdir], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)] + (['multi'] if multi else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    f = r.stdout.split()
    out_h, out_w, tiles, launches, hit = int(f[-5]), int(f[-4]), int(f[-3]), int(f[-2]), float(f[-1])
    assert (out_h, out_w, tiles) == (362, 362, 144) and launches > 0
    assert hit > 0.85, r.stdout          # img1(x) = img2(x - 3): elevation = j - (j - 3) = 3 wherever the match lies inside the tile (not in its 3 leftmost columns)
