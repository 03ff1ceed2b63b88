"""The tile-level algorithm of the planned fused MRF stage (tools/fused_mrf_proto.py) against the oracle's generator:
halo recompute + zero-masked intermediates reproduce the layer-wise result; without the mask they do not."""
import os, sys
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle.voice_loader import load_voice
from oracle.vits_oracle import Oracle
from piper_b200 import voicegen
import fused_mrf_proto as fp


@pytest.mark.parametrize("arch", ["tiny", "tiny-high"])       # ResBlock2 and ResBlock1 generators
@pytest.mark.parametrize("TO", [184, 64, 1000])
def test_fused_tiles_match_layerwise(arch, TO):
    spec, w, attrs = load_voice(voicegen.cached_voice(arch))
    o = Oracle(spec, w, attrs)
    ids = voicegen.benchmark_ids(6, seed=3)
    dump = {}
    o.infer(ids, (0.667, 1.0, 0.8), dump=dump)
    tw = {k: torch.as_tensor(v) for k, v in w.items() if k.startswith("dec.resblocks")}
    for stage in range(len(spec.up_rates)):
        x, ref = dump[f"up{stage}"], dump[f"stage{stage}"]
        chains = fp.resblock_chains(tw, spec, stage)
        got = fp.fused_stage(x, chains, TO=TO)
        assert got.shape == ref.shape
        scale = float(ref.abs().max())
        assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, scale), (arch, stage)
        # the padding pitfall: unmasked intermediates differ at the utterance edges (and only there)
        bad = fp.fused_stage(x, chains, TO=TO, mask=False)
        err = (bad - ref).abs().amax(0)
        H = max(fp.halo(c) for c in chains)
        assert float(err.max()) > 1e-4 * max(1.0, scale)
        if x.shape[1] > 2 * H + 2:
            assert float(err[H:x.shape[1] - H].max()) <= 2e-5 * max(1.0, scale)


def test_halos_of_the_piper_presets():
    class S: pass
    mk = lambda k: torch.zeros(4, 4, k)
    # medium / x-low / low: ResBlock2, kernels (3,5,7), dilations (1,2),(2,6),(3,12)  
    chains = [[(mk(3), None, 1, 0), (mk(3), None, 2, 1)], [(mk(5), None, 2, 0), (mk(5), None, 6, 1)],
              [(mk(7), None, 3, 0), (mk(7), None, 12, 1)]]
    assert [fp.halo(c) for c in chains] == [3, 16, 45]
    # high: ResBlock1, kernels (3,7,11), c1 dilations (1,3,5), c2 dilation 1
    hi = []
    for k in (3, 7, 11):
        c = []
        for d in (1, 3, 5):
            c += [(mk(k), None, d, None), (mk(k), None, 1, len(c))]
        hi.append(c)
    assert [fp.halo(c) for c in hi] == [12, 36, 60]
