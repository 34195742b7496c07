"""Where the end-to-end path (host threads calling Matching(data) with pinned host tensors) loses against the device-timed
throughput: per-thread host time inside forward vs time blocked in the result read-back, for 1 .. 16 caller threads."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import Matching
from gims_b200.synth import make_pair, make_state_dict
n = 2048
dev = torch.device('cuda:0')
m = Matching({})
m.gmodel.load_state_dict(make_state_dict(0))
m = m.eval().to(dev)
pool = []
for s in range(16):
    d = make_pair(n, n, seed=100 + s)
    d = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in d.items()}
    d['device'] = dev
    pool.append(d)

pool_dev = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()} for d in pool]


def run(T, per_thread, read_back=True, stagger_ms=0.0, resident=False, raw=False, pinned_sync=False):
    src = pool_dev if resident else pool
    streams = [torch.cuda.Stream(device=dev) for _ in range(T)]
    pins = [torch.empty(8 + 2 * 4096, dtype=torch.int32).pin_memory() for _ in range(T)]
    t_fwd = [0.0] * T
    t_read = [0.0] * T
    def worker(t):
        torch.cuda.set_device(dev)
        if stagger_ms:
            time.sleep(t * stagger_ms * 1e-3)
        with torch.no_grad(), torch.cuda.stream(streams[t]):
            for i in range(per_thread):
                a = time.perf_counter()
                d = src[(t + i) % 16]
                if raw:      # the library call + the one metadata read-back, none of the dict handling of Matching.forward
                    r = m.gmodel.run_pair(d['keypoints0'][0], d['descriptors0'][0], d['scores0'][0], d['keypoints1'][0],
                                          d['descriptors1'][0], d['scores1'][0], d['image0'].shape, d['image1'].shape)
                    if pinned_sync:
                        buf = pins[t][:r['meta'].numel()]
                        buf.copy_(r['meta'], non_blocking=True)
                        streams[t].synchronize()
                    else:
                        r['meta'].cpu()
                    pred = None
                else:
                    pred = m(dict(d))
                b = time.perf_counter()
                if read_back and pred is not None:
                    pred['matches0'].cpu(); pred['matching_scores0'].cpu()
                c = time.perf_counter()
                t_fwd[t] += b - a; t_read[t] += c - b
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    t0 = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    k = T * per_thread
    print('T=%2d read_back=%d resident=%d raw=%d pinned_sync=%d: %.1f pairs/s | per pair and thread: forward() %.2f ms, read-back %.2f ms' %
          (T, read_back, resident, raw, pinned_sync, k / dt, 1e3 * sum(t_fwd) / k, 1e3 * sum(t_read) / k))

run(8, 4)
for rep in range(2):
    run(8, 64, raw=True)
    run(8, 64, raw=True, pinned_sync=True)
    run(12, 48, raw=True, pinned_sync=True)
    run(16, 32, raw=True, pinned_sync=True)
