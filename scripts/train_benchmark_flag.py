"""Experiment: training step time with torch.backends.cudnn.benchmark off / on (python scripts/train_benchmark_flag.py 0|1).
Measured on B200, config 2, B = 16: 64.2 ms vs 58.4 ms per step."""
import sys, os, argparse, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.backends.cudnn.benchmark = (sys.argv[1] == "1")
from mde_biological_vision_systems_b200 import synthetic
from mde_biological_vision_systems_b200.ExternalInfoLoaders.SemanticsLoader import SemanticsLoader
dev = torch.device("cuda:0")
args = argparse.Namespace(steps=5, batch=16)
B, H, W = 16, 416, 544
loader = SemanticsLoader(argparse.Namespace(use_semantics=bench.SEM_MODE), device=dev)
host = {"image": synthetic.image(B, H, W, seed=0).pin_memory(), "depth": synthetic.depth(B, H, W, seed=1).pin_memory(),
        "semantics": synthetic.label_maps(B, H, W, seed=2)[0].pin_memory()}
r = bench.run_train(args, dev, 1, 0, host, loader)
print("cudnn.benchmark", sys.argv[1], r["ms_per_step"], r["value"], r["loss"])
