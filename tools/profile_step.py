"""One denoising iteration of the 8B path (image step: cond forward + text step + uncond forward + image step) for
ncu. Only the region between cudaProfilerStart/Stop is profiled (run ncu with --profile-from-start off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mmada_parallel_b200.generators.parallel_generator import DenoiseState, denoise_loop
from mmada_parallel_b200.schedule import cosine_schedule

layers = int(os.environ.get("MMDP_PROFILE_LAYERS", "32"))
cfg = dict(bench.MODEL_8B, n_layers=layers)
model = bench.build_model(cfg, "cuda:0", seed=1000)
lay = bench.synthetic_layout(0)
pos_args = {k: lay[k] for k in ("text_start", "text_end", "image_start", "seq_len", "newline_every")}
rng = torch.Generator(device="cuda:0").manual_seed(42)
kw = dict(temperature=1.0, text_temperature=0.0, cfg_scale=0.0, cfg_img=4.0, noise_schedule=cosine_schedule,
          text_vocab_size=bench.TEXT_VOCAB, codebook_size=bench.CODEBOOK, generator=rng)
def state():
    return DenoiseState(model, lay["input_ids"], uncon_text=lay["uncon_text"], uncon_image=lay["uncon_image"], cfg_scale=0.0,
                        cfg_img=4.0, codebook_size=bench.CODEBOOK, **pos_args)
with torch.no_grad():
    denoise_loop(state(), text_steps=4, timesteps=4, **kw)   # warm-up: 4 iterations, 3 of them image steps
    torch.cuda.synchronize()
    st = state()
    torch.cuda.profiler.start()
    denoise_loop(st, text_steps=1, timesteps=1, **kw)        # exactly one iteration, and it is an image step
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one image-step iteration with", layers, "layers")
