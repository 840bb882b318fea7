import sys, time, threading, torch
sys.path.insert(0, '/root/repo')
import dips_b200, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
w, hgt, n, fmt = 1920, 1080, 1800, 0
fb = w*hgt*3
clip = torch.empty((n, fb), dtype=torch.uint8, device='cuda')
dips_b200.synth_fill_device(0, clip.data_ptr(), 0, n, w, hgt, fmt, 1, 1, 0)
torch.cuda.synchronize()
ctx = dips_b200.Context(w, hgt, fmt, 0, 32)
samples = []
stop = False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetPowerUsage(h)/1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.02)
for name in ("probe", "clip"):
    samples.clear(); stop = False
    t = threading.Thread(target=sampler); t.start()
    t0 = time.time()
    if name == "probe":
        ms = ctx.stream_probe(clip.data_ptr(), n, fb, reps=900)
    else:
        ctx.enable_timing(True); ctx.clip_kernel_time()
        for _ in range(900):
            ctx.reset(); ctx.run_clip_device(clip.data_ptr(), n)
        tot, cnt = ctx.clip_kernel_time(); ms = tot/cnt
    dt = time.time() - t0
    stop = True; t.join()
    tail = samples[len(samples)//2:]
    print(name, "kernel ms", round(ms, 4), "GB/s", round(n*fb/ms/1e6), "wall", round(dt, 2), "power W (2nd half avg/max)", round(sum(p for p, _ in tail)/len(tail)), round(max(p for p, _ in tail)), "sm MHz (2nd half avg)", round(sum(c for _, c in tail)/len(tail)))
    time.sleep(3)
