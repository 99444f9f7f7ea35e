import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from vad_b200 import runtime
from vad_b200.analyser import StreamBank
h = runtime.Handle(0, ffn_weights=runtime.glorot_ffn(0))
bank = StreamBank(4096, handle=h)
rng = np.random.default_rng(0)
pool = torch.from_numpy((rng.standard_normal((32, 4096, 160)) * 3000).astype(np.int16))
def run(n, fn):
    lat = []
    for t in range(n + 100):
        bank.h_chunks.copy_(pool[t % 32])
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        if t >= 100: lat.append(dt)
    ms = np.sort(np.array(lat)) * 1e3
    return round(float(ms[len(ms)//2]), 4), round(float(ms[int(len(ms)*0.99)]), 4)
print("feed_pinned", run(3000, bank.feed_pinned))
# kernel alone, device time
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(200):
    bank.bank.feed_ptr(bank.d_chunks.data_ptr(), bank.d_labels.data_ptr(), 0)
ev1.record(); torch.cuda.synchronize()
print("kernel_us", ev0.elapsed_time(ev1) / 200 * 1e3)
g = bank._graphs[False]
def raw():
    g.replay(); torch.cuda.synchronize()
print("replay+devsync", run(3000, raw))
st = torch.cuda.current_stream()
def raw2():
    g.replay(); st.synchronize()
print("replay+streamsync", run(3000, raw2))
bank3 = StreamBank(4096, handle=h, zero_copy=False)
def run3(n, fn):
    lat = []
    for t in range(n + 100):
        bank3.h_chunks.copy_(pool[t % 32])
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        if t >= 100: lat.append(dt)
    ms = np.sort(np.array(lat)) * 1e3
    return round(float(ms[len(ms)//2]), 4), round(float(ms[int(len(ms)*0.99)]), 4)
print("copy-node graph", run3(3000, bank3.feed_pinned))
