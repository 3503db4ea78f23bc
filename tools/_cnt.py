import torch, sys
sys.path.insert(0,".")
import open_pi_zero_b200 as pz
from open_pi_zero_b200.pizero import PiZeroInference
from open_pi_zero_b200.synth import fill_random_
d = pz.make_dims(vocab_size=1024, image_token_index=1000, num_layers=2, vit_layers=2)
m = PiZeroInference(pz.cfg_from_dims(d), init="empty", device="cuda", dtype=torch.bfloat16)
fill_random_(m, d); m.use_cuda_graph=False
B=int(sys.argv[1])
inp = pz.make_inputs(d, B, seed=0)
out = m(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda().bfloat16(), proprios=inp["proprios"].cuda(), noise=inp["noise"].cuda(), valid_len=inp["valid_len"].cuda())
torch.cuda.synchronize()
print(B, m.last_launch_count)
