"""CTA-pair issue-rate probe (test infrastructure): cycles per tcgen05.mma.cta_group::2 (M = 256 over two SMs, B
split N/2 + N/2) next to the single-CTA M = 128 instruction of the same N, K-major SW64 fp16 operands with
distinct tiles per MMA (the conv main loop's pattern).  -> gpurun_out/probe_pair.json"""
import ctypes
import json
import os

import numpy as np
import torch

import probe_umma as pu
import probe_rate as pr

lib = pu.lib
lib.probe_mma_rate_pair.restype = ctypes.c_int
lib.probe_mma_rate_pair.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                    ctypes.c_void_p, ctypes.c_void_p]
OUT = []


def rate_pair(N, nacc, repeat=256, rowb=64, a_sbo=None, a_offs=None, b_offs=None, dcols=None, label="pair_M256"):
    rng = np.random.default_rng(0)
    mode = pu.SWZ_64B if rowb == 64 else pu.SWZ_128B
    rows = 1024 if rowb == 64 else 512
    X = (0.01 * rng.standard_normal((rows, rowb // 2))).astype(np.float16)
    A_OFF, B_OFF = 0, ((rows * rowb + 1023) // 1024) * 1024
    img = pu.Image(2 * B_OFF)
    img.put_rows(A_OFF, X.view(np.uint8).reshape(rows, rowb), mode)
    img.put_rows(B_OFF, X.view(np.uint8).reshape(rows, rowb), mode)
    adesc = pu.make_desc(A_OFF, 16, a_sbo if a_sbo else 8 * rowb, mode)
    bdesc = pu.make_desc(B_OFF, 16, 8 * rowb, mode)
    nt = max(1, min(8, rows // 128))
    nb = max(1, min(8, rows // max(N // 2, 8)))
    ao = np.array([(j % nt) * (128 * rowb >> 4) for j in range(8)], dtype=np.uint32)
    bo = np.array([(j % nb) * ((N // 2) * rowb >> 4) for j in range(8)], dtype=np.uint32)
    if a_offs is not None:
        ao = np.array(a_offs, dtype=np.uint32)
    if b_offs is not None:
        bo = np.array(b_offs, dtype=np.uint32)
    idesc = pu.make_idesc(pu.FMT_F16, 256, N, 0, 0)
    dev = torch.device("cuda")
    img_t = torch.from_numpy(img.buf).to(dev)
    cyc = torch.zeros(4, dtype=torch.int64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    res = {}
    for rep in (8, repeat):
        rc = lib.probe_mma_rate_pair(img_t.data_ptr(), img.buf.size, adesc, bdesc, idesc, ao.ctypes.data, bo.ctypes.data,
                                     nacc, dcols if dcols else N, rep, cyc.data_ptr(), st.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.probe_last_error().decode())
        c = cyc.cpu().numpy()
        res[rep] = (int(c[0]), int(c[1]))
    per_total = (res[repeat][0] - res[8][0]) / (8.0 * (repeat - 8))
    r = dict(label=label, N=N, rowb=rowb, nacc=nacc, cyc_per_mma=round(per_total, 2),
             macs_per_cyc_per_sm=round(128 * N * 16 / per_total, 1), frac_of_4096=round(128 * N * 16 / per_total / 4096, 3),
             status=int(st.item()))
    OUT.append(r)
    print(json.dumps(r), flush=True)


def main():
    print(torch.cuda.get_device_name(0))
    for rowb in (64,):
        for N in (128, 160, 256):
            for nacc in (1, 2):
                if nacc * N <= 512:
                    rate_pair(N, nacc, rowb=rowb)
    # conv-like operand walks (conv_pair.cu): A = haloed brick rows (12 voxels per h row -> SBO 768 B), start shifted by
    # the (kh, kw) tap and the K half; B = 5120-byte weight half-stages
    taps = [(kh * 12 + kw) * 64 >> 4 for kh in range(5) for kw in range(5)]
    plane = 15360 >> 4
    rate_pair(160, 2, label="conv_A_sbo768_only", a_sbo=768, a_offs=[0] * 8)
    rate_pair(160, 2, label="conv_A_taps", a_sbo=768, a_offs=[taps[(3 * j) % 25] for j in range(8)])
    rate_pair(160, 2, label="conv_A_taps_khalf", a_sbo=768, a_offs=[taps[(3 * (j // 2)) % 25] + 2 * (j % 2) for j in range(8)])
    rate_pair(160, 2, label="conv_A_planes_taps_khalf", a_sbo=768,
              a_offs=[(j // 2) * plane + taps[7] + 2 * (j % 2) for j in range(8)])
    rate_pair(160, 2, label="conv_B_khalf", b_offs=[(j // 2 % 6) * (5120 >> 4) + 2 * (j % 2) for j in range(8)])
    rate_pair(160, 2, label="conv_like_all", a_sbo=768, a_offs=[(j // 2) * plane + taps[7] + 2 * (j % 2) for j in range(8)],
              b_offs=[2 * (j % 2) for j in range(8)])
    rate_pair(160, 2, label="conv_like_all_dcol32", a_sbo=768, dcols=32,
              a_offs=[(j // 2) * plane + taps[7] + 2 * (j % 2) for j in range(8)], b_offs=[2 * (j % 2) for j in range(8)])
    rate_pair(160, 1, label="conv_like_same_acc", a_sbo=768,
              a_offs=[(j // 2) * plane + taps[7] + 2 * (j % 2) for j in range(8)], b_offs=[2 * (j % 2) for j in range(8)])
    # does a tcgen05.commit between MMAs cost tensor-pipe time?
    lib.probe_set_commit_every.restype = ctypes.c_int
    lib.probe_set_commit_every.argtypes = [ctypes.c_int]
    for ce in (8, 4, 2, 1):
        lib.probe_set_commit_every(ce)
        rate_pair(160, 2, label=f"pair_commit_every_{ce}")
    lib.probe_set_commit_every(0)
    # single-CTA reference points from the same build
    for N in (128, 160, 256):
        nt = 8 if N <= 128 else 1024 // N
        pr.rate(0, N, False, 64, 2 if 2 * N <= 512 else 1, a_offs=[(j % 8) * (128 * 64 >> 4) for j in range(8)],
                b_offs=[(j % nt) * (N * 64 >> 4) for j in range(8)], label="single_M128")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_pair.json", "w") as f:
        json.dump(OUT + pr.OUT, f, indent=1)


if __name__ == "__main__":
    main()
