"""Worker for test_sharded_sampling_world_size_2_gloo (launched by torch.distributed.run, gloo, CPU)."""
import sys

import torch
import torch.distributed as dist

from diffusion_conductor_b200.generate import sharded_sample
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict
from oracle import motion_oracle as O


def main():
    dist.init_process_group("gloo")
    torch.set_num_threads(2)
    sd = synth_state_dict(5, num_layers=1)
    B, T = 5, 12
    xf_proj, xf_out = synth_features(B, T, seed=2)
    _, noise = synth_inputs(B, T, seed=2)
    length = [12, 3, 12, 7, 1]
    tb = O.Tables(O.linear_betas(25))

    def sample_fn(xp, xo, nz, ln):      # CPU stand-in for the CUDA sampler: the oracle
        return O.sample_loop(sd, tb, nz, ln, xp, xo, max_steps=2)[0]

    gathered = sharded_sample(sample_fn, xf_proj, xf_out, noise, length)        # 5 clips over 2 ranks: unequal shards (padded gather)
    single = sample_fn(xf_proj, xf_out, noise, length)
    even = sharded_sample(sample_fn, xf_proj[:4], xf_out[:4], noise[:4], length[:4])   # equal shards: one all_gather_into_tensor
    torch.save({"gathered": gathered, "single": single, "even": even}, f"{sys.argv[1]}.{dist.get_rank()}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
