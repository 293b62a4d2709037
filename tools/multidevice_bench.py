#!/usr/bin/env python3
"""Frame sharding INSIDE the library (qldpc_decoder_config.devices[]), one process: BASELINE config 2 through the
host-pointer call `qldpc_decode_bits` with pinned buffers, 65 536 frames per device, for 1 .. all visible devices.
This is what a single-process `ecd2` gets; `bench.py` under torchrun is the one-process-per-GPU counterpart.

    python tools/multidevice_bench.py [--out gpurun_out/multidevice.json]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-device", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    q = importlib.import_module("qcrypto-ldpc_b200")
    L = q.lib()
    code = q.Code.from_qc_file(q.data_path("NR_1_1_384.qc"))
    N, K = code.n, code.k
    ndev = torch.cuda.device_count()
    results = []
    for n in [d for d in (1, 2, 4, 8) if d <= ndev]:
        F = args.frames_per_device * n
        dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True,
                        norm_factor=0.75, out_mode=q.OUT_INFO, devices=list(range(n)) if n > 1 else None, flags=q.FLAG_L2_PERSIST)
        # synthetic frames: encode on device 0 in pieces, BSC on the information bits, pinned host buffers
        one = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, norm_factor=0.75,
                        out_mode=q.OUT_INFO, device=0)
        h_bits = torch.empty((F, dec.cw_words), dtype=torch.int32).pin_memory()
        h_out = torch.empty((F, dec.out_words), dtype=torch.int32).pin_memory()
        h_ok = torch.empty(F, dtype=torch.uint8).pin_memory()
        h_it = torch.empty(F, dtype=torch.int16).pin_memory()
        g = torch.Generator(device="cuda:0"); g.manual_seed(3)
        msg_all = torch.empty((F, K // 32), dtype=torch.int32)
        st = torch.cuda.current_stream(0).cuda_stream
        for f0 in range(0, F, 65536):
            nf = min(65536, F - f0)
            msg = torch.randint(-2**31, 2**31 - 1, (nf, K // 32), dtype=torch.int64, device="cuda:0", generator=g).to(torch.int32)
            cw = torch.empty((nf, dec.cw_words), dtype=torch.int32, device="cuda:0")
            one.encode_nr_device(msg.data_ptr(), nf, cw.data_ptr(), st)
            flip = (torch.rand((nf, K), device="cuda:0", generator=g) < 0.03)
            w = (2 ** torch.arange(31, -1, -1, device="cuda:0", dtype=torch.int64))
            fw = (flip.view(nf, -1, 32).to(torch.int64) * w).sum(dim=-1)
            fw = torch.where(fw >= 2**31, fw - 2**32, fw).to(torch.int32)
            cw[:, :K // 32] ^= fw
            torch.cuda.synchronize()
            h_bits[f0:f0 + nf].copy_(cw)
            msg_all[f0:f0 + nf].copy_(msg)
        known = np.zeros(N, np.uint8); known[K:] = 1
        kn = np.ascontiguousarray(q.pack_bits(known[None, :])[0])
        one.close()

        def call():
            rc = L.qldpc_decode_bits(dec.h, h_bits.data_ptr(), kn.ctypes.data, None, 14.0, 31.0, None, F, h_out.data_ptr(),
                                     h_ok.data_ptr(), h_it.data_ptr())
            assert rc == 0, rc
        for _ in range(3):
            call()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            call()
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        ok = bool(h_ok.all()) and bool((h_out[:, :K // 32] == msg_all).all())
        res = {"n_devices": n, "frames": F, "ms_per_call": ms, "e2e_info_mbps": F * K / ms / 1e3, "all_reconciled": ok,
               "api": "qldpc_decode_bits, one process, devices[] in the decoder config, pinned host buffers"}
        print(json.dumps(res), flush=True)
        results.append(res)
        dec.close()
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
