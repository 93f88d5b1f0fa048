"""Discrete-event model (not a measurement) of conv_chain_kernel's schedule against layer-by-layer launches; unit times are
rough fits to profiles/r1c_conv_counters_timeline.txt.  usage: python tools/chain_sim.py"""
# Discrete-event model of conv_chain_kernel's scheduling: 74 clusters, units in list order round-robin, a unit's MMA loop may
# start when (a) the cluster's previous MMA loop is done and (b) the 3x3 tile neighbourhoods of its source layers are complete
# (epilogues done).  Epilogues are serial per cluster and overlap the next MMA loop (double-buffered TMEM).
import math
B, tiles_y, tiles_x = 32, 2, 5
tpi = tiles_x * tiles_y; m_tiles = B * tpi; m_groups = (m_tiles + 1) // 2
NCL = 74
CLK = 1.965e3  # clocks per us
# layer: (name, taps, chunks, cout_pad, n_tile, epi_clocks_per_unit, srcs, halo)
layers = [("C1", 1, 6, 256, 256, 12000, [], 0), ("C2", 9, 4, 192, 192, 9000, [0], 1), ("F1", 1, 2, 128, 128, 6000, [], 0),
          ("F2", 9, 2, 64, 64, 4000, [2], 1), ("ENC", 9, 4, 128, 128, 12000, [1, 3], 1), ("ZR1", 5, 4, 256, 256, 16000, [4], 1),
          ("Q1", 5, 4, 128, 128, 14000, [5], 1), ("ZR2", 5, 4, 256, 256, 16000, [6], 1), ("Q2", 5, 4, 128, 128, 14000, [7], 1),
          ("HEADS", 9, 2, 512, 256, 15000, [8], 1), ("MASK2", 1, 4, 576, 192, 10000, [9], 0)]
def mma_clocks(taps, chunks, n_tile):        # 12 MMAs per stage, N/2 clocks each (M=256 pair), + ~150 clocks of issue overhead
    return taps * chunks * (12 * max(n_tile / 2, 49) + 150)
units = []          # (layer, n_idx, group)
for l, (name, taps, chunks, cout_pad, n_tile, epi, srcs, halo) in enumerate(layers):
    for n_idx in range(cout_pad // n_tile):
        for g in range(m_groups):
            units.append((l, n_idx, g))
def nbrs(t, halo):
    b, r = divmod(t, tpi); ty, tx = divmod(r, tiles_x)
    if not halo: return [t]
    return [b * tpi + y * tiles_x + x for y in range(max(0, ty - 1), min(tiles_y, ty + 2)) for x in range(max(0, tx - 1), min(tiles_x, tx + 2))]
def simulate(chained):
    done_t = {}                                  # (layer, tile) -> time all its n-units' epilogues are complete
    cnt = {}
    mma_free = [0.0] * NCL; epi_free = [0.0] * NCL
    layer_end = [0.0] * len(layers)
    barrier = 0.0
    pending = {c: [u for i, u in enumerate(units) if i % NCL == c] for c in range(NCL)}
    ptr = {c: 0 for c in range(NCL)}
    finish = 0.0
    # process units in global order (each cluster's order is a subsequence; deps always point backwards)
    for i, (l, n_idx, g) in enumerate(units):
        c = i % NCL
        name, taps, chunks, cout_pad, n_tile, epi, srcs, halo = layers[l]
        tiles = [t for t in (2 * g, 2 * g + 1) if t < m_tiles]
        ready = 0.0
        if chained:
            for sl in srcs:
                for t in tiles:
                    for nb in nbrs(t, halo):
                        ready = max(ready, done_t[(sl, nb)])
        else:
            ready = max([layer_end[sl] for sl in range(l)] + [0.0]) + (8000 if l else 0)    # grid-wide wait + launch/fill gap
        start = max(mma_free[c], ready)
        mma_end = start + mma_clocks(taps, chunks, n_tile)
        mma_free[c] = mma_end
        e_end = max(mma_end, epi_free[c]) + epi
        epi_free[c] = e_end
        for t in tiles:
            k = (l, t); cnt[k] = cnt.get(k, 0) + 1
            if cnt[k] == cout_pad // n_tile: done_t[k] = e_end
            else: done_t[k] = max(done_t.get(k, 0.0), e_end)
        layer_end[l] = max(layer_end[l], e_end)
        finish = max(finish, e_end)
    return finish / CLK
print("layer-by-layer launches (model): %.0f us" % simulate(False))
print("chained single launch (model):   %.0f us" % simulate(True))
ideal = sum(mma_clocks(t, c, n) * (cp // n) * m_groups for (_, t, c, cp, n, _, _, _) in layers) / NCL / CLK
print("sum of MMA loops / 74 clusters:   %.0f us" % ideal)
