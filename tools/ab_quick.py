import ctypes, os, sys, torch
lib = os.environ.get("SHB_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seqhasher_b200", "libseqhash_b200.so")
L = ctypes.CDLL(lib)
c = ctypes
L.shb_init.argtypes = [c.c_int, c.c_int]
L.shb_hash_device.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_size_t, c.c_size_t, c.POINTER(c.c_int), c.c_int, c.c_uint, c.POINTER(c.c_void_p), c.c_void_p, c.c_void_p, c.c_void_p]
L.shb_synth_fill.argtypes = [c.c_int, c.c_void_p, c.c_size_t, c.c_uint64, c.c_uint, c.c_uint, c.c_void_p]
L.shb_scratch_bytes.argtypes = [c.c_size_t, c.c_size_t]; L.shb_scratch_bytes.restype = c.c_size_t
L.shb_init(1, 0)
n, Ln = 10_000_000, 250
total = n * Ln
st = torch.cuda.current_stream().cuda_stream
d_res = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
L.shb_synth_fill(0, d_res.data_ptr(), total, 0xC2, 50, 0, st)
d_offs = torch.arange(0, n + 1, dtype=torch.int64, device="cuda") * Ln
scr = torch.empty(L.shb_scratch_bytes(total, n), dtype=torch.uint8, device="cuda")
sizes = {0: 20, 4: 8, 2: 16, 3: 8}
for algo in (0, 4, 2, 3):
    for flags in (1, 3):
        out = torch.empty(n * sizes[algo], dtype=torch.uint8, device="cuda")
        ids = (c.c_int * 1)(algo); ptrs = (c.c_void_p * 1)(out.data_ptr())
        run = lambda: L.shb_hash_device(0, d_res.data_ptr(), d_offs.data_ptr(), n, total, ids, 1, flags, ptrs, None, scr.data_ptr(), st)
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        print(os.path.basename(lib), "algo", algo, "flags", flags, round(e0.elapsed_time(e1) / 10, 4), "ms", flush=True)
