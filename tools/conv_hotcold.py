"""Hot (weights resident in L2) vs cold (distinct weights per launch, > L2 in total) timing of the tcgen05 convolution."""
import sys
import torch
sys.path.insert(0, ".")
from tools.conv_bench import LAYERS, time_layer
dev = torch.device("cuda:0")
for idx in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else (10, 0):
    spec = LAYERS[idx]
    for engine in ("umma", "tf32"):
        hot, fl = time_layer(dev, spec, engine, nlayers=1, reps=50)
        cold, _ = time_layer(dev, spec, engine, nlayers=12, reps=10)
        print(f"{spec[0]:40s} {engine:5s} hot {hot:7.1f} us ({fl / hot / 1e6:6.1f} TF/s)   cold {cold:7.1f} us ({fl / cold / 1e6:6.1f} TF/s)", flush=True)
