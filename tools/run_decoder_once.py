"""Runs the bench workload's inference path once with a short decoder loop (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tacotron2_b200 as t2
from tacotron2_b200 import _capi
from bench import synth_weights, B_PER_GPU, T_TEXT
from tests.common import rand_text

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
model = t2.Tacotron2(t2.create_hparams()); model.load_state_dict(synth_weights()); model = model.cuda().eval()
model.decoder.max_decoder_steps = steps; model.decoder.gate_threshold = 1.0
eng = model._t2_engine(); eng.impl = _capi.IMPL_PERSISTENT
text = rand_text(B_PER_GPU, T_TEXT, 100).cuda()
with torch.no_grad():
    for _ in range(2):
        out = model.inference(text)
torch.cuda.synchronize()
print("ok", out[0].shape)
