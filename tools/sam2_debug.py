"""Stage-by-stage comparison of the CUDA SAM2 forward with transformers' Sam2Model (CPU fp32) on a GPU box."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from PIL import Image
from oracle import sam2_hf, thumbnail as ot
from atlaspatch_b200.sam2 import B200Sam2Predictor
from atlaspatch_b200.synthetic import make_spec, render_region_host

torch.set_num_threads(16)
sd = sam2_hf.sam2_state_dict(0)
model = sam2_hf.build_model(sd)
spec = make_spec(8192, 8192, 0)
thumb = ot.area_reduce(render_region_host(spec, 0, 0, 8192, 8192), 16)
img = np.array(Image.fromarray(thumb).resize((1024, 1024), Image.Resampling.BILINEAR))

acts = {}
def hook(name):
    def f(mod, inp, out):
        acts[name] = (out[0] if isinstance(out, tuple) else out).detach()
    return f
bb = model.vision_encoder.backbone
bb.patch_embed.register_forward_hook(hook("pe_raw"))
for i, blk in enumerate(bb.blocks):
    blk.register_forward_hook(hook(f"blk{i}"))
model.vision_encoder.neck.register_forward_hook(lambda m, i, o: acts.__setitem__("fpn", o[0]))
model.mask_decoder.transformer.register_forward_hook(lambda m, i, o: acts.__setitem__("dec", o))
t0 = time.time()
up_ref, low_ref = sam2_hf.predict_logits(model, img)
print("HF forward %.2fs" % (time.time() - t0))

pred = B200Sam2Predictor(sd)
torch.cuda.synchronize() if torch.cuda.is_available() else None
t0 = time.time(); up, low = pred.predict_logits(img, return_lowres=True); t1 = time.time()
up, low = pred.predict_logits(img, return_lowres=True); t2 = time.time()
print("B200 forward first %.3fs second %.3fs" % (t1 - t0, t2 - t1))

def cmp(name, got, ref):
    ref = ref.reshape(got.shape)
    err = np.abs(got - ref).max(); sc = np.abs(ref).max()
    rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    print(f"{name:12s} max|diff| {err:.3e}  max|ref| {sc:.3e}  rel-l2 {rel:.3e}")

pe = (acts["pe_raw"][0] + bb._get_pos_embed((256, 256))[0]).detach()
cmp("patch_embed", pred.debug_buffer("patch_embed", (256 * 256, 96)), pe.reshape(-1, 96).numpy())
for i, blk in enumerate(bb.blocks):
    ref = acts[f"blk{i}"][0].numpy()
    cmp(f"blk{i}", pred.debug_buffer(f"blk{i}", (ref.shape[0] * ref.shape[1], ref.shape[2])), ref)
fpn = acts["fpn"]   # tuple in neck order: [lvl3, lvl2, lvl1, lvl0], NCHW
for lvl, t in zip((3, 2, 1, 0), fpn):
    ref = t[0].permute(1, 2, 0).numpy()
    cmp(f"fpn{lvl}", pred.debug_buffer(f"fpn{lvl}", (ref.shape[0] * ref.shape[1], 256)), ref)
q_ref, k_ref = acts["dec"]
cmp("queries", pred.debug_buffer("queries", (9, 256)), q_ref[0, 0].numpy())
cmp("keys", pred.debug_buffer("keys", (4096, 256)), k_ref[0, 0].numpy())
cmp("low_res", low, low_ref)
cmp("logits", up, up_ref)
a, b = up > 0, up_ref > 0
print("mask IoU %.6f  positives %d vs %d  disagreeing pixels %d" % ((a & b).sum() / max((a | b).sum(), 1), a.sum(), b.sum(), (a ^ b).sum()))
