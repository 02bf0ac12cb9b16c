"""Where does tcgen05.mma put the accumulator rows for M = 64 (cta_group::1)?  Uses tools/umma_probe.cu (K-major SW128 operands, N = 64,
K = 64) and matches every TMEM lane's 64 columns against the rows of A @ B^T.  (development tool)
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC -o tools/libumma_probe.so tools/umma_probe.cu"""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import umma_probe as up

rng = np.random.default_rng(1)
A = rng.integers(-4, 5, (64, 64)).astype(np.float32)
B = rng.integers(-4, 5, (64, 64)).astype(np.float32)
e = 2
a_img = torch.from_numpy(up.image(A, up.kmajor_sw(2), e, False, 96 * 1024)).cuda()
b_img = torch.from_numpy(up.image(B, up.kmajor_sw(2), e, False, 96 * 1024)).cuda()
D = torch.full((128, 64), float('nan'), device='cuda')
rc = up.lib.probe_run(a_img.data_ptr(), 96 * 1024, b_img.data_ptr(), 96 * 1024, 0, 0, up.desc(0, 1024, up.SW128), up.desc(0, 1024, up.SW128),
                      up.idesc(64, 64), 4, 32, 32, 0, 64, D.data_ptr(), None)
torch.cuda.synchronize()
ref = A.astype(np.float64) @ B.astype(np.float64).T
got = D.cpu().numpy().astype(np.float64)
lane_of_row = {}
for lane in range(128):
    for i in range(64):
        if np.array_equal(got[lane], ref[i]):
            lane_of_row.setdefault(i, []).append(lane)
print(json.dumps(dict(rc=rc, rows_found=len(lane_of_row), lane_of_row={str(k): v for k, v in sorted(lane_of_row.items())})))
