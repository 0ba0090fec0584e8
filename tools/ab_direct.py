"""A/B of the hash kernels on chosen length distributions (SHB_LIB selects the library build, so two builds can be
compared in one gpurun call).  Prints ms per pass, residue GB/s and a digest checksum per (shape, algorithm, flags).
usage: python tools/ab_direct.py [shapes] [algorithms] [flags]
  shapes      comma list of SHAPES keys (default: all)          e.g. readme,mid1k,c4big
  algorithms  comma list of names (default: all but xxh3/nthash) e.g. sha1,md5,xxh3
  flags       comma list of SHB_FLAG_* values (default 0,1)      e.g. 1,3  or  0,0x200 (0x200 = no long-record kernels,
              0x100 = force the direct path, 0x400 = no length sort)
Standalone ctypes loader so that older builds of the library can be compared."""
import ctypes, os, sys, torch
c = ctypes
lib = os.environ.get("SHB_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seqhasher_b200", "libseqhash_b200.so")
L = ctypes.CDLL(lib)
L.shb_init.argtypes = [c.c_int, c.c_int]
L.shb_hash_device.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_size_t, c.c_size_t, c.POINTER(c.c_int), c.c_int, c.c_uint, c.POINTER(c.c_void_p), c.c_void_p, c.c_void_p, c.c_void_p]
L.shb_synth_fill.argtypes = [c.c_int, c.c_void_p, c.c_size_t, c.c_uint64, c.c_uint, c.c_uint, c.c_void_p]
L.shb_scratch_bytes.argtypes = [c.c_size_t, c.c_size_t]; L.shb_scratch_bytes.restype = c.c_size_t
L.shb_init(1, 0)
st = torch.cuda.current_stream().cuda_stream
SIZES = {0: 20, 1: 64, 2: 16, 3: 8, 4: 8, 5: 16, 6: 16, 7: 8, 8: 32, 9: 128, 10: 8}
NAMES = {0: "sha1", 1: "sha3", 2: "md5", 3: "xxhash", 4: "xxh3", 5: "cityhash", 6: "murmur3", 7: "nthash", 8: "blake3", 9: "k12", 10: "rapidhash"}
SHAPES = {"readme": (500_000, 30, 3000), "mid1k": (1_000_000, 800, 1200), "c4": (20_000, 10_000, 50_000),
          "u500_1500": (700_000, 500, 1500), "u100_2000": (700_000, 100, 2000), "u900_2100": (500_000, 900, 2100), "fix1000": (700_000, 1000, 1000), "c4big": (150_000, 10_000, 50_000), "amp": (5_000_000, 100, 600), "amp250": (5_000_000, 200, 300), "fix250": (5_000_000, 250, 250), "u2k_8k": (300_000, 2000, 8000), "u5k_15k": (150_000, 5000, 15000), "u8k_24k": (100_000, 8000, 24000), "u50k_100k": (40_000, 50_000, 100_000), "fix5000": (200_000, 5000, 5000), "u100k_200k": (20_000, 100_000, 200_000)}
only = sys.argv[1].split(",") if len(sys.argv) > 1 else list(SHAPES)
algo_filter = sys.argv[2].split(",") if len(sys.argv) > 2 else [v for v in NAMES.values() if v not in ("xxh3", "nthash")]
FLAGS = [int(x, 0) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]   # 0x200 = SHB_FLAG_NO_LONG_PATH
g = torch.Generator(device="cpu"); g.manual_seed(7)
for shape in only:
    n, lo, hi = SHAPES[shape]
    lens = torch.randint(lo, hi + 1, (n,), generator=g, dtype=torch.int64)
    offs = torch.zeros(n + 1, dtype=torch.int64); offs[1:] = torch.cumsum(lens, 0)
    total = int(offs[-1])
    d_offs = offs.cuda()
    d_res = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
    L.shb_synth_fill(0, d_res.data_ptr(), total, 0xAB, 50, 0, st)
    scr = torch.empty(L.shb_scratch_bytes(total, n), dtype=torch.uint8, device="cuda")
    for algo in SIZES:
        if NAMES[algo] not in algo_filter:
            continue
        for flags in FLAGS:
            out = torch.zeros(n * SIZES[algo], dtype=torch.uint8, device="cuda")
            ids = (c.c_int * 1)(algo); ptrs = (c.c_void_p * 1)(out.data_ptr())
            run = lambda: L.shb_hash_device(0, d_res.data_ptr(), d_offs.data_ptr(), n, total, ids, 1, flags, ptrs, None, scr.data_ptr(), st)
            for _ in range(3): assert run() == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            chk = int(out.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFF   # digests must agree across builds
            print(f"{os.path.basename(lib)} {shape} {NAMES[algo]} flags={flags} {ms:.4f} ms {total / ms / 1e6:.0f} GB/s chk={chk:012x}", flush=True)
