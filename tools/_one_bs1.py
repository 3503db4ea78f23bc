import os, sys
sys.path.insert(0, "/root/repo")
import torch
import open_pi_zero_b200 as pz
from open_pi_zero_b200.pizero import PiZeroInference
from open_pi_zero_b200.synth import fill_random_
dims = pz.make_dims(num_layers=3, vit_layers=3)
dev = torch.device("cuda")
m = PiZeroInference(pz.cfg_from_dims(dims), init="empty", device=dev, dtype=torch.bfloat16)
fill_random_(m, dims)
m.use_cuda_graph = False
inp = pz.make_inputs(dims, 1, seed=0)
kw = dict(input_ids=inp["input_ids"].to(dev), pixel_values=inp["pixel_values"].to(dev, torch.bfloat16), proprios=inp["proprios"].to(dev), noise=inp["noise"].to(dev), valid_len=inp["valid_len"].to(dev))
for _ in range(3):
    m(**kw)
torch.cuda.synchronize()
