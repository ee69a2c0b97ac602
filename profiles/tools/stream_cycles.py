"""How fast an SM streams L2-resident weight tiles into shared memory: cp.async.bulk vs 16-byte cp.async (csrc/spl_umma.cu:
spl_umma_stream_cycles), by ring depth, tile size, number of issuing warps, pieces per tile and number of CTAs pulling at once."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(__file__), "..", "..")))
import azg_b200
from azg_b200 import _native as nat
lib = nat.lib(); dev = torch.device("cuda", 0)
h = C.c_void_p(); nat.check(lib.spl_ctx_create(2, 10, nat.RULES_DEFAULT, 0, C.byref(h)))
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
src = torch.zeros(36 * 16384, dtype=torch.uint8, device=dev)
def run(grid, kind, nw, pieces, tile_bytes, depth):
    out = torch.zeros(grid, dtype=torch.int64, device=dev)
    src_tiles = 36 * 16384 // tile_bytes
    tiles = 4 * src_tiles
    mode = kind + 16 * (nw - 1) + 256 * (pieces - 1)
    for _ in range(2):
        nat.check(lib.spl_umma_stream_cycles(h, C.c_void_p(src.data_ptr()), src_tiles, tile_bytes, depth, tiles, mode, grid, C.c_void_p(out.data_ptr()), st))
        torch.cuda.synchronize()
    cyc = float(out.max().item())
    print(f"grid {grid:3d} {('bulk', 'cp.async16', 'bulk, a lane per slot')[kind]} warps {nw} pieces {pieces} tile {tile_bytes:5d} B depth {depth}: {cyc / tiles:8.1f} cycles per tile = "
          f"{tile_bytes * tiles / cyc:6.1f} B/clk per SM, {grid * tile_bytes * tiles / cyc * 1.965:8.0f} GB/s chip", flush=True)
for grid in (1, 148):
    run(grid, 2, 1, 1, 16384, 6)
    run(grid, 2, 1, 1, 16384, 12)
    run(grid, 2, 1, 1, 8192, 12)
    run(grid, 2, 2, 1, 16384, 12)
    run(grid, 2, 1, 1, 65536, 2)
    run(grid, 0, 1, 1, 16384, 6)
    run(grid, 0, 1, 1, 65536, 2)
    run(grid, 0, 1, 1, 32768, 4)
    run(grid, 0, 1, 4, 16384, 6)
    run(grid, 0, 1, 16, 16384, 6)
    run(grid, 0, 2, 1, 16384, 6)
    run(grid, 0, 4, 1, 16384, 8)
    run(grid, 0, 8, 1, 16384, 8)
    run(grid, 0, 4, 4, 16384, 8)
    run(grid, 1, 1, 1, 16384, 6)
    run(grid, 1, 2, 1, 16384, 8)
    run(grid, 1, 4, 1, 16384, 8)
    run(grid, 1, 8, 1, 16384, 8)
    run(grid, 1, 4, 1, 4096, 8)
