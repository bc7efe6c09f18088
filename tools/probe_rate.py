"""Issue-rate probe (test infrastructure): cycles per tcgen05.mma for the instruction shapes the conv
kernels can use, with 1/2/4/8 independent TMEM accumulators round-robin. -> gpurun_out/probe_rate.json"""
import ctypes
import json
import os

import numpy as np
import torch

import probe_umma as pu

lib = pu.lib
lib.probe_mma_rate.restype = ctypes.c_int
lib.probe_mma_rate.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                               ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32,
                               ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
OUT = []


def rate(kind, N, mn_major, rowb, nacc, repeat=256, a_offs=None, b_offs=None, label=""):
    rng = np.random.default_rng(0)
    mode = pu.SWZ_64B if rowb == 64 else pu.SWZ_128B
    esz = 2 if kind == 0 else 4
    chans = rowb // esz
    rows = 1024 if rowb == 64 else 512       # two regions (A, B) of 64 KB each
    if kind == 0:
        X = (0.01 * rng.standard_normal((rows, chans))).astype(np.float16)
    else:
        X = pu.tf32_round((0.01 * rng.standard_normal((rows, chans))).astype(np.float32))
    A_OFF, B_OFF = 0, ((rows * rowb + 1023) // 1024) * 1024
    img = pu.Image(2 * B_OFF)
    img.put_rows(A_OFF, X.view(np.uint8).reshape(rows, rowb), mode)
    img.put_rows(B_OFF, X.view(np.uint8).reshape(rows, rowb), mode)
    n_k = rowb // 32
    if not mn_major:
        adesc = pu.make_desc(A_OFF, 16, 8 * rowb, mode)
        bdesc = pu.make_desc(B_OFF, 16, 8 * rowb, mode)
        ao = a_offs if a_offs is not None else [(j % n_k) * 2 for j in range(8)]
        bo = b_offs if b_offs is not None else [(j % n_k) * 2 for j in range(8)]
    else:
        adesc = pu.make_desc(A_OFF, rowb, 8 * rowb, mode)
        bdesc = pu.make_desc(B_OFF, rowb, 8 * rowb, mode)
        ao = a_offs if a_offs is not None else [j * (16 * rowb >> 4) for j in range(8)]
        bo = b_offs if b_offs is not None else [j * (16 * rowb >> 4) for j in range(8)]
    idesc = pu.make_idesc(pu.FMT_F16 if kind == 0 else pu.FMT_TF32, 128, N, int(mn_major), int(mn_major))
    dev = torch.device("cuda")
    img_t = torch.from_numpy(img.buf).to(dev)
    cyc = torch.zeros(4, dtype=torch.int64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    ao_a = np.array(ao, dtype=np.uint32)
    bo_a = np.array(bo, dtype=np.uint32)
    res = {}
    for rep in (8, repeat):
        rc = lib.probe_mma_rate(img_t.data_ptr(), img.buf.size, adesc, bdesc, idesc, kind, ao_a.ctypes.data,
                                bo_a.ctypes.data, nacc, N, rep, cyc.data_ptr(), st.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.probe_last_error().decode())
        c = cyc.cpu().numpy()
        res[rep] = (int(c[0]), int(c[1]))
    per_total = (res[repeat][0] - res[8][0]) / (8.0 * (repeat - 8))
    per_issue = (res[repeat][1] - res[8][1]) / (8.0 * (repeat - 8))
    kk = 16 if kind == 0 else 8
    r = dict(label=label, kind=kind, N=N, mn_major=mn_major, rowb=rowb, nacc=nacc, cyc_per_mma=round(per_total, 2),
             issue_cyc_per_mma=round(per_issue, 2), macs_per_cyc=round(128 * N * kk / per_total, 1),
             frac_of_4096=round(128 * N * kk / per_total / (4096 if kind == 0 else 2048), 3), status=int(st.item()))
    OUT.append(r)
    print(json.dumps(r), flush=True)


def main():
    print(torch.cuda.get_device_name(0))
    for kind, rowb in ((0, 128), (0, 64), (1, 128)):
        for N in (32, 64, 128, 256):
            for nacc in (1, 2, 4, 8):
                if nacc * N <= 512:
                    rate(kind, N, False, rowb, nacc, label="kmajor")
    for N in (32, 128, 160, 256):
        for nacc in (1, 2, 4):
            if nacc * N <= 512:
                rate(0, N, True, 64, nacc, label="mnmajor")
    # conv-like: A start shifted by whole rows (tap shifts), B walks distinct weight tiles
    for N, nacc in ((32, 8), (64, 8), (128, 4)):
        rate(0, N, False, 64, nacc, a_offs=[(j % 5) * 4 + (j % 2) * 2 for j in range(8)],
             b_offs=[j * (N * 64 >> 4) // 4 * 4 % 2048 for j in range(8)], label="convlike_sw64")
    # fully distinct operand tiles per MMA (as in the real conv main loop): A tile j = rows [128j, 128j+128),
    # B tile j = rows [Nj, Nj+N)  -> does the operand fetch cost depend on re-touching the same bytes?
    for kind, rowb in ((0, 64), (0, 128)):
        for N in (32, 96, 128, 160, 256):
            nt = max(1, min(8, (1024 if rowb == 64 else 512) // max(N, 128)))
            rate(kind, N, False, rowb, 4 if N * 4 <= 512 else (2 if N * 2 <= 512 else 1),
                 a_offs=[(j % nt) * (128 * rowb >> 4) for j in range(8)],
                 b_offs=[(j % nt) * (N * rowb >> 4) for j in range(8)], label="distinct_tiles")
    # same A for consecutive MMAs, distinct B (A re-use), and vice versa
    for N in (128, 160):
        nb = 1024 // N
        rate(0, N, False, 64, 2, a_offs=[0] * 8, b_offs=[(j % nb) * (N * 64 >> 4) for j in range(8)], label="sameA_distinctB")
        rate(0, N, False, 64, 2, a_offs=[(j % 8) * (128 * 64 >> 4) for j in range(8)], b_offs=[0] * 8, label="distinctA_sameB")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_rate.json", "w") as f:
        json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()
