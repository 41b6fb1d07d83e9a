"""Where the time of bench.py's pack pass goes: every K1 launch of the Llama-3-8B conversion alone (DeviceModel of
bench.py), per launch and for 1 .. 256 back-to-back passes, with the SM clock read through NVML."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awq-converter_b200"))
import torch
import bench
from awq_quantizer import _native as N
from awq_quantizer import model_shapes as M
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
specs = [(n, s, ck) for n, s, ck in M.workload("llama3-8b") if M.numel(s) >= 128]
model = bench.DeviceModel(torch, N, M, specs, dev, g=128, sym=False, T=2048, n_grid=20, search=True)
model.grids(); model.search(final=False); torch.cuda.synchronize()
bpe = 2 + 0.5 + 2.0 / 128 + 0.5 / 128

def ev_time(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = []
for mode in ("searched scales", "unit scales"):
    if mode == "unit scales":
        for n, *_ in model.searched:
            model.sel[n][2].fill_(1.0)
    for i in range(0, len(model.searched), model.FINAL_CHUNK):
        chunk = model.searched[i:i + model.FINAL_CHUNK]
        el = sum(C * K for _, C, K, _ in chunk)
        model.finals(chunk)
        ms = min(ev_time(lambda: model.finals(chunk), 3) for _ in range(3))
        out.append({"mode": mode, "chunk": i // model.FINAL_CHUNK, "tensors": len(chunk), "ms": round(ms, 4),
                    "frac": round(bpe * el / (ms * 1e-3) / 1e9 / 6546.6, 3)})
        print(json.dumps(out[-1]), flush=True)
el = sum(model.w[n].numel() for n, *_ in model.plain)
ms = min(ev_time(model.k1_plain, 3) for _ in range(3))
print(json.dumps({"arena": True, "ms": round(ms, 4), "frac": round(bpe * el / (ms * 1e-3) / 1e9 / 6546.6, 3)}), flush=True)
for reps in (1, 2, 4, 16, 64, 256):
    ms = ev_time(model.pack_only, reps)
    clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
    out.append({"passes": reps, "ms_per_pass": round(ms, 4), "frac": round(bpe * model.elems / (ms * 1e-3) / 1e9 / 6546.6, 3),
                "sm_mhz_after": clk, "power_w_after": pw})
    print(json.dumps(out[-1]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_pack_pass.json"), "w"), indent=1)
