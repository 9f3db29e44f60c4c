"""Variant timer: python profiles/tools/time_variant.py <libmsort.so> [kind] — state/obs checksums and us/step of one
build of the library (MSORT_LIB), used to A/B kernel variants on the same box."""
import os, sys, hashlib
lib = sys.argv[1]; kind = sys.argv[2] if len(sys.argv) > 2 else "mono"
os.environ["MSORT_LIB"] = lib
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
n = int(os.environ.get("N", 1 << 20))
cls = {"mono": ms.BatchedMonolithEnv, "sort": ms.BatchedSortingEnv, "press": ms.BatchedPressingEnv}[kind]
env = cls(n, max_steps=50, seed=42, info_level=os.environ.get("INFO", "episode"))
if kind == 'press' and not os.environ.get('NOPOLICY'):
    from marl_sortingenv_b200.policy import sb3_style_init
    env.set_sort_policy(sb3_style_init(0, action_gain=float(os.environ.get("GAIN", 0.01))))
    if os.environ.get("TENSOR") is not None:                      # A/B: embedded policy on the tensor cores (1) or FFMA2 (0)
        env.set_option(ms._abi.OPT_TENSOR_POLICY, int(os.environ["TENSOR"]))
    if os.environ.get("CTAS"):
        env.set_option(ms._abi.OPT_PERSIST_CTAS, int(os.environ["CTAS"]))
    print("persistent CTAs per SM:", env.get_option(ms._abi.OPT_PERSIST_CTAS), flush=True)
env.reset()
if os.environ.get("L2PERSIST"):
    # experiment: pin the hot state planes in L2 with an access-policy window on the stream (captured into the graph's kernel nodes)
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    class Win(C.Structure):
        _fields_ = [("base_ptr", C.c_void_p), ("num_bytes", C.c_size_t), ("hitRatio", C.c_float), ("hitProp", C.c_int), ("missProp", C.c_int)]
    class Attr(C.Union):
        _fields_ = [("win", Win), ("pad", C.c_char * 64)]
    hot = 64 * ((n + 127) // 128 * 128)
    print("set limit rc", rt.cudaDeviceSetLimit(C.c_int(6), C.c_size_t(min(hot, 96 << 20))))
    a = Attr(); a.win = Win(env.state.data_ptr(), hot, float(os.environ["L2PERSIST"]), 2, 1)   # hit: persisting, miss: streaming
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    print("set attr rc", rt.cudaStreamSetAttribute(C.c_void_p(st.cuda_stream), C.c_int(1), C.byref(a)))
T = 128
acts = torch.zeros((T, n), dtype=torch.int64, device="cuda")
rsum = torch.zeros((), dtype=torch.float64, device="cuda")
for t in range(T):
    env.sample_actions(7, t, out=acts[t]); env.step(acts[t]); rsum += env.reward.double().sum()
torch.cuda.synchronize()
h = hashlib.sha1(env.state.cpu().numpy().tobytes()).hexdigest()[:12]
ho = hashlib.sha1(env.obs.cpu().numpy().tobytes()).hexdigest()[:12]
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for t in range(512):
        env.step(acts[t % T])
g.replay(); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 512 * 1e3)
print(f"{os.path.basename(lib):28s} {kind:5s} n {n} [{env.step_variant}] state {h} obs {ho} rsum {rsum.item():.6f}  us/step {best:.2f}  G/s {n/best/1e3:.2f}", flush=True)
