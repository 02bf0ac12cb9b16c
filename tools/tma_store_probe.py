"""Drive tools/tma_store_probe.cu on a B200: which 4-D TMA store forms work (alignment of the shared-memory source,
negative / out-of-bounds coordinates, box wider than the tensor).  One case per subprocess.
    python tools/tma_store_probe.py            -> gpurun_out/tma_store_probe.json"""
import ctypes, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (Cin, Cout, W, H, N, load box w, smem_off, load coords, store coords, nrep, rep_dy, store box w, store skip bytes)
    'aligned_inbounds':        (64, 64, 8, 8, 2, 8, 0, (0, 0, 1, 0), (0, 0, 1, 0), 1, 0, 8, 0),
    'negx_store':              (64, 64, 8, 8, 2, 10, 0, (0, -1, 1, 0), (0, -1, 1, 0), 1, 0, 10, 0),
    'skip_border_src128':      (64, 64, 8, 8, 2, 10, 0, (0, -1, 1, 0), (0, 0, 1, 0), 1, 0, 8, 128),
    'skip_border_off_rows3':   (64, 64, 8, 8, 2, 10, 1280, (0, -1, 1, 0), (0, 0, 1, 0), 3, 1, 8, 128),
    'clip_upper_x':            (64, 64, 8, 8, 2, 10, 0, (0, 0, 1, 0), (0, 0, 1, 0), 1, 0, 10, 0),
    'clip_upper_y':            (64, 64, 8, 8, 2, 8, 0, (0, 0, 7, 0), (0, 0, 8, 0), 1, 0, 8, 0),
    'narrowC32':               (64, 32, 8, 8, 2, 10, 0, (0, -1, 1, 0), (0, 0, 1, 0), 1, 0, 8, 128),
    'narrowC32_off_rows3':     (64, 32, 8, 8, 2, 10, 1280, (0, -1, 1, 0), (0, 0, 1, 0), 3, 1, 8, 128),
    'W32_rows3_off_c64':       (128, 128, 32, 32, 3, 34, 4352, (64, -1, 29, 1), (64, 0, 29, 1), 3, 1, 32, 128),
}


def run_case(name):
    import torch
    Cin, Cout, W, H, N, boxw, off, lc, sc, nrep, dy, boxs, skip = CASES[name]
    lib = ctypes.CDLL(os.path.join(HERE, 'libtma_store_probe.so'))
    lib.tma_store_probe.restype = ctypes.c_int
    lib.tma_store_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_long] * 2 + [ctypes.c_int] * 14
    g = torch.Generator(); g.manual_seed(1)
    ld_in, ld_out = Cin + 8, Cout + 24
    x = torch.randn(N * H * W, ld_in, generator=g).to(torch.bfloat16).cuda()
    out = torch.full((N * H * W, ld_out), -7.0, dtype=torch.bfloat16, device='cuda')
    rc = lib.tma_store_probe(x.data_ptr(), out.data_ptr(), Cin, Cout, W, H, N, ld_in, ld_out, boxw, off, *lc, *sc, nrep, dy, boxs, skip)
    if rc:
        return dict(case=name, ok=False, rc=rc)
    exp = torch.full_like(out, -7.0)
    xv, ev = x.view(N, H, W, ld_in), exp.view(N, H, W, ld_out)
    for r in range(nrep):
        n, y = sc[3], sc[2] + r * dy
        if not (0 <= n < N and 0 <= y < H):
            continue
        for i in range(boxs):
            xx = sc[1] + i
            if 0 <= xx < W:
                for c in range(64):
                    ci, co = lc[0] + c, sc[0] + c
                    if co < Cout:
                        ev[n, y, xx, co] = xv[lc[3], lc[2] + r * dy, lc[1] + i + skip // 128, ci] if (ci < Cin and 0 <= lc[2] + r * dy < H and 0 <= lc[1] + i + skip // 128 < W) else 0
    bad = int((out != exp).sum())
    return dict(case=name, ok=bad == 0, rc=0, mismatches=bad)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        print(json.dumps(run_case(sys.argv[1])))
        sys.exit(0)
    res = []
    for name in CASES:
        p = subprocess.run([sys.executable, __file__, name], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        line = [l for l in p.stdout.splitlines() if l.startswith('{')]
        r = json.loads(line[-1]) if line else dict(case=name, ok=False, rc=p.returncode, err=p.stderr[-300:])
        res.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(HERE, '..', 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(HERE, '..', 'gpurun_out', 'tma_store_probe.json'), 'w'), indent=1)
