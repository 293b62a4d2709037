import importlib, math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
q = importlib.import_module("qcrypto-ldpc_b200")
dev = torch.device("cuda", 0)
code = q.Code.from_qc_file(q.data_path("qkd_psdpeg_n65536.qc"))
N = code.n
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
g = torch.Generator(device=dev); g.manual_seed(1)
for rule, name, norm in ((q.RULE_SPA, "SPA", 1.0), (q.RULE_NMS, "NMS", 0.8125)):
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=rule, dtype=q.DTYPE_F32, max_iter=50, early_stop=True, norm_factor=norm, out_mode=q.OUT_ALL)
    st = torch.cuda.current_stream().cuda_stream
    qber = 0.03
    x = torch.randint(0, 2, (F, N), dtype=torch.uint8, device=dev, generator=g)
    e = (torch.rand((F, N), device=dev, generator=g) < qber).to(torch.uint8)
    w = (2 ** torch.arange(31, -1, -1, device=dev, dtype=torch.int64))
    def pack(b):
        v = (b.view(F, -1, 32).to(torch.int64) * w).sum(dim=-1)
        return torch.where(v >= 2**31, v - 2**32, v).to(torch.int32).contiguous()
    xb, yb = pack(x), pack(x ^ e)
    syn = torch.empty((F, dec.syn_words), dtype=torch.int32, device=dev)
    dec.syndrome_device(xb.data_ptr(), F, syn.data_ptr(), st)
    llr = torch.empty((F, N), dtype=torch.float32, device=dev)
    dec.make_llr_device(yb.data_ptr(), 0, 0, math.log((1 - qber) / qber), 0.0, F, llr.data_ptr(), st)
    out = torch.empty((F, dec.cw_words), dtype=torch.int32, device=dev); ok = torch.empty(F, dtype=torch.uint8, device=dev); it = torch.empty(F, dtype=torch.int16, device=dev)
    for _ in range(2):
        dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dec.decode_device(llr.data_ptr(), syn.data_ptr(), F, out.data_ptr(), ok.data_ptr(), it.data_ptr(), 0, st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(dec.kernel_name, name, "F", F, "ok", bool(ok.all()), "iters", float(it.float().mean()), "ms", round(dt * 1e3, 2), "Mbit/s", round(F * N / dt / 1e6))
