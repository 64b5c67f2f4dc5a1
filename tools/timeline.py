"""Print the CTA-0 event timeline of the layer kernel for one denoise step of C2 (run under gpurun)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer, _lib  # noqa: E402
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict  # noqa: E402

NAMES = {1: "start", 2: "rows_done"}
ROWW = ["qA_sa", "S_sa", "Wo_sa(H)", "Wq_ca", "qA_ca", "S_ca", "Wo_ca(H)", "W1", "W2", "S_ff", "Wo_ff(H)", "QKV", "KtV pass", "KtV pass2", "?"]
ROWP = ["film_sa", "ln_ca", "softmax_ca", "film_ca", "h_bf16", "gelu", "film_ff", "ln_sa", "E,V images", "V image 2", "?"]
DOPS = ["qA_sa", "Wo_sa", "Wq_ca", "qA_ca", "Wo_ca", "W1", "W2", "Wo_ff", "Wq", "Wk", "Wv", "KtV", "?", "?"]
B, T, S = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 180, 50)))
m = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device="cuda", music_model_path=None)
m.load_state_dict(synth_state_dict(0), strict=True)
m = m.cuda().eval()
d = GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                      model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
xf_proj, xf_out = synth_features(B, T, seed=1)
_, noise = synth_inputs(B, T, seed=1)
x = noise.cuda()
eng = d._bind(m, x, dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B))
for _ in range(3):
    eng.sample_step(_lib.DC_SAMPLER_DDIM, x.clone(), S - 1, None)
torch.cuda.synchronize()
nl = 9
buf = (C.c_uint64 * (nl * 512))()
xs = x.clone()
_lib.check(eng.lib.dc_debug_timeline(eng.handle, C.c_void_p(xs.data_ptr()), S - 1, buf, nl), eng.handle)
persistent = T <= 2048 and os.environ.get("DC_PERSIST", "1") != "0"
if persistent:          # cluster-per-clip kernel: three lanes (row thread 0, dependent-MMA issuer, FiLM issuer) of (clock, id) pairs
    LANE = 1 + 2 * 680
    ev = []
    for lane_id in range(3):
        cnt = min(int(buf[lane_id * LANE]), 680)
        ev += [(int(buf[lane_id * LANE + 1 + 2 * i]), int(buf[lane_id * LANE + 2 + 2 * i])) for i in range(cnt)]
    ev.sort()
    n = len(ev)
    t0 = ev[0][0]
    names = {1: "start", 2: "rows_done", 112: "rows: got QKV", 208: "        mma: issued QKV", 209: "        mma: issued KtV pass"}
    dn = ["qA_sa", "Wo_sa", "Wq_ca", "qA_ca", "Wo_ca", "W1", "W2", "Wo_ff"]
    lo, hi = (int(v) for v in os.environ.get("TL_RANGE", "0,140").split(","))
    print(f"--- persistent clip kernel: {n} events; showing events [{lo},{hi}) ; total {ev[-1][0] - t0} cycles")
    for k, (t, i) in enumerate(ev):
        if not (lo <= k < hi):
            continue
        if i in names:
            nm = names[i]
        elif 120 <= i < 130:
            nm = "rows epi: " + ["col max/E/V done", "-", "published V", "got KtV", "partial written", "all partials visible", "published merged image", "A_emb image written", "prologue operands staged", "h0 in TMEM"][i - 120]
        elif 100 < i < 150:
            nm = "rows: got " + ROWW[i - 101]
        elif 150 < i < 200:
            nm = "rows: published " + ROWP[i - 151]
        elif 200 <= i < 208:
            nm = "        mma: issued " + dn[i - 200]
        elif i == 400:
            nm = "        mma: a_ready observed"
        elif i == 401:
            nm = "        mma: ring-B stage full"
        elif 310 <= i < 320:
            nm = f"                mma2: S#{i - 310} last stage issued"
        else:
            nm = str(i)
        print(f"{t - t0:8d}  {nm}")
    sys.exit(0)
for launch in (1, 4):
    seg = buf[launch * 512:(launch + 1) * 512]
    n = min(int(seg[0]), 254)
    ev = sorted((int(seg[1 + 2 * i]), int(seg[2 + 2 * i])) for i in range(n))
    t0 = ev[0][0]
    print(f"--- launch {launch} (layer {launch - 1}): {n} events, cycles relative to start")
    for t, i in ev:
        if i in NAMES:
            nm = NAMES[i]
        elif 120 <= i < 130:
            nm = "rows epi: " + ["col max done", "E image + sums done", "published V", "got KtV", "partials written", "counters done"][i - 120]
        elif 100 < i < 150:
            nm = "rows: got " + ROWW[i - 101] if launch > 0 else f"rows wait {i}"
        elif 150 < i < 200:
            nm = "rows: published " + ROWP[i - 151]
        elif 200 <= i < 300:
            nm = "        mma: issued " + DOPS[i - 200]
        elif 300 <= i < 310:
            nm = f"        mma: S#{i - 300} first stage"
        else:
            nm = f"        mma: S#{i - 310} last stage issued"
        print(f"{t - t0:8d}  {nm}")
