"""CPU, build container only: run the unmodified reference model.py live against the oracle on
fresh random inputs (skipped where /root/reference is absent, e.g. on the GPU box)."""
import pytest
import torch

from oracle import tacotron2_oracle as O
from oracle.ref_import import (MaskInjector, default_hparams, import_reference_model,
                               injected_dropout, reference_available)
from tests.common import keep_mask, rand_text, rel_err, synth_state_dict

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


def test_reference_inference_b1_live():
    ref = import_reference_model()
    sd = synth_state_dict(5, gate_bias=-10.0, scale=2.0)
    model = ref.Tacotron2(default_hparams()); model.load_state_dict(sd); model.eval()
    model.decoder.max_decoder_steps = 12
    text = rand_text(1, 19, 3); keep = keep_mask((12, 2, 1, 256), 0.5, 4)
    masks = [keep[t, l].bool() for t in range(12) for l in range(2)]
    with torch.no_grad(), injected_dropout(ref, MaskInjector(masks)):
        r = model.inference(text)
    with torch.no_grad():
        mel, post, gate, align, lengths = O.tacotron2_inference(sd, text, keep, 0.5, 12)
    assert int(lengths[0]) == r[0].shape[2] == 12
    for a, b in zip((mel, post, gate, align), r):
        assert rel_err(a, b) < 2e-5


def test_reference_state_dict_layout():
    """The 84 keys / shapes the boundary must reproduce."""
    from tests.common import state_dict_shapes
    ref = import_reference_model()
    sd = ref.Tacotron2(default_hparams()).state_dict()
    want = state_dict_shapes()
    assert list(sd.keys()) == list(want.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(want[k]), k
