"""Ad-hoc parity soak: random image shapes / feature counts / seeds, CUDA extractor against the oracle (bit-exact).
Usage (GPU box): python tools/soak_extract.py [cases] [seed]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import orb_slam2_chinesenotes_b200 as ob
from oracle_lib import OracleExtractor
from synth import synth_frame

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for c in range(cases):
    w, h = int(rng.integers(120, 1400)), int(rng.integers(100, 1000))
    nf = int(rng.integers(100, 3000))
    nlev = int(rng.choice([4, 8, 8, 8, 10]))
    sf = float(rng.choice([1.2, 1.2, 1.15, 1.3]))
    ini, mn = int(rng.choice([20, 20, 30, 12])), int(rng.choice([7, 7, 5, 10]))
    seed = int(rng.integers(0, 1 << 30))
    try:
        G = ob.ORBextractor(nf, sf, nlev, ini, mn)
    except ob.OrbError:
        continue
    if not G._L.orbx_shape_supported(G._h, w, h):
        G.close(); continue
    O = OracleExtractor(nf, sf, nlev, ini, mn)
    img = synth_frame(w, h, seed)
    if c % 5 == 0:
        img = (img.astype(np.int32) // 32 * 32).astype(np.uint8)          # plateaus: score ties everywhere
    n, k_o, d_o = O.extract(img)
    k_g, d_g = G(img, capacity=max(2 * nf, n + 64))
    ok = n == len(k_g) and all((k_o[f] == k_g[f]).all() for f in k_o.dtype.names) and (d_o == d_g).all()
    print(c, (w, h, nf, nlev, sf, ini, mn), n, "ok" if ok else "MISMATCH", flush=True)
    bad += 0 if ok else 1
    G.close(); O.close()
print("mismatches:", bad)
