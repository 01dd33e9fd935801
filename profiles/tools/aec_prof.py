import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import aec_streams as S, evk_loader
evk = evk_loader.load(); evk.lib()
case = sys.argv[1] if len(sys.argv) > 1 else "quiet"
e = S.stream(34, 10000, blobs=3, noise=0.02, tie=50) if case == "quiet" else S.stream(31, 10000, tie=1250)
with evk.Evk(1024) as h:
    h.aec_create(None)
    for i in range(0, len(e), 1250):
        h.aec_update(e[i:i+1250])
