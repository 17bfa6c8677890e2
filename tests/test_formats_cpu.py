"""CPU tests of videoswap_b200/formats.py (the data formats either side of the path) against
tests/golden/formats.pt, which oracle/make_golden_formats.py produced with the reference's OWN functions
(merge_lora_into_weight, select_frame_idx, bind_concept_prompt)."""
import hashlib
import os

import pytest
import torch

from videoswap_b200 import formats as F

GOLD = os.path.join(os.path.dirname(__file__), "golden", "formats.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


class _ParamBag(torch.nn.Module):
    """A module whose named_parameters() are the given state dict under its own key names (what the merge walks)."""

    def __init__(self, sd):
        super().__init__()
        self._names = {}
        for i, (k, v) in enumerate(sd.items()):
            self.register_parameter(f"p{i}", torch.nn.Parameter(v.clone(), requires_grad=False))
            self._names[f"p{i}"] = k
        self.dirty = 0

    def named_parameters(self, *a, **k):
        for n, p in super().named_parameters(*a, **k):
            yield self._names[n], p

    def mark_weights_dirty(self):
        self.dirty += 1


def _tiny_sd(gold, dtype):
    from videoswap_b200.spec import UNetConfig, unet_param_shapes
    from videoswap_b200.weights import seeded_state_dict
    t = gold["tiny"]
    cfg = UNetConfig(block_out_channels=t["boc"], cross_attention_dim=t["ctx"], norm_num_groups=t["groups"])
    return {k: v.to(dtype) for k, v in seeded_state_dict(unet_param_shapes(cfg), seed=0).items()}


def _digest(sd, keys):
    h = hashlib.sha256()
    for k in sorted(keys):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("tag,dtype", [("fp32", torch.float32), ("fp16", torch.float16)])
def test_edlora_merge_is_bit_exact_with_the_reference_and_restores(gold, tag, dtype):
    sd = _tiny_sd(gold, dtype)
    bag = _ParamBag(sd)
    backup = F.merge_edlora_into_unet(bag, gold["lora"], gold["alpha"])
    now = dict(bag.named_parameters())
    assert sorted(backup) == gold[tag]["touched"] and bag.dirty == 1
    for k, ref in gold[tag]["sample"].items():
        assert now[k].dtype == dtype and torch.equal(now[k], ref), k
    assert _digest(now, backup.keys()) == gold[tag]["digest"]
    untouched = [k for k in sd if k not in backup]
    assert all(torch.equal(now[k], sd[k]) for k in untouched)
    F.restore_unet(bag, backup)
    assert bag.dirty == 2 and all(torch.equal(p, sd[k]) for k, p in bag.named_parameters())


def test_edlora_merge_strict_rejects_stray_tensors(gold):
    bag = _ParamBag(_tiny_sd(gold, torch.float32))
    with pytest.raises(KeyError, match="match no UNet weight"):
        F.merge_edlora_into_unet(bag, gold["lora"], 0.6, strict=True)
    lora = {k: v for k, v in gold["lora"].items() if "not_a_unet_key" not in k}
    down = next(k for k in lora if "lora_down" in k)
    broken = {k: v for k, v in lora.items() if k != down}
    with pytest.raises(KeyError, match="but not"):
        F.merge_edlora_into_unet(_ParamBag(_tiny_sd(gold, torch.float32)), broken, 0.6)


def test_lora_down_name_follows_the_reference_replace_chain():
    pre = "up_blocks.1.attentions.0.transformer_blocks.0."
    assert F.lora_down_name(pre + "attn2.to_out.0.weight") == pre + "attn2.to_out.0.lora_down.weight"
    assert F.lora_down_name(pre + "ff.net.0.proj.weight") == pre + "ff.net.0.proj.lora_down.weight"
    assert F.lora_down_name("mid_block.attentions.0.proj_in.weight") == "mid_block.attentions.0.proj_in.lora_down.weight"
    assert F.lora_down_name(pre + "attn1.to_q.bias") == pre + "attn1.to_q.bias"          # biases / norms / convs: no LoRA site
    assert F.lora_down_name("conv_in.weight") == "conv_in.weight"


def test_load_edlora_file_layouts(tmp_path, gold):
    emb = {"<cat>": torch.randn(16, 768)}
    p = tmp_path / "edlora.pth"
    torch.save({"params": {"new_concept_embedding": emb, "unet": gold["lora"], "text_encoder": {}}}, p)
    d = F.load_edlora(p)
    assert set(d) == {"new_concept_embedding", "unet", "text_encoder"} and set(d["unet"]) == set(gold["lora"])
    assert F.load_edlora({"unet": gold["lora"]})["new_concept_embedding"] == {}            # without the 'params' wrapper
    names = F.new_concept_token_names(emb)
    assert names == {"<cat>": [f"<<cat>_{i}>" for i in range(16)]}
    assert F.new_concept_token_names(emb, enable_edlora=False) == {"<cat>": ["<<cat>_0>"]}


def test_select_frame_idx_and_bind_concept_prompt_match_the_reference(gold):
    for (b, e, n), ref in gold["select_frame_idx"]:
        assert F.select_frame_idx(b, e, n) == ref
    with pytest.raises(ValueError):
        F.select_frame_idx(0, 16, 1)
    for prompts, cfg, ref in gold["bind"]:
        assert F.bind_concept_prompt(prompts, cfg) == ref
        assert F.bind_concept_prompt(prompts, {k: v["concept_token_names"] for k, v in cfg.items()}) == ref


def test_tap_file_to_conditions(tmp_path):
    g = torch.Generator().manual_seed(0)
    tracks = 512 * torch.rand((40, 5, 2), generator=g)
    tracks[3, 2] = -1
    tap = {"pred_tracks": tracks, "point_name2id": {"nose": 0, "tail": 1, "paw_l": 2, "paw_r": 3, "ear": 4},
           "point_embedding": torch.randn((5, 1280), generator=g)}
    p = tmp_path / "TAP.pth"
    torch.save(tap, p)
    sel = F.select_frame_idx(0, 40, 16)
    c = F.load_tap(p, select_id=sel, img_size=(768, 448))
    assert c["pred_tracks"].shape == (16, 5, 2) and torch.equal(c["pred_tracks"], tracks[sel])
    assert c["img_size"] == (768, 448) and c["index_list"] is None and torch.equal(c["point_embedding"], tap["point_embedding"])
    c2 = F.load_tap(p, select_id=sel, img_size=(768, 448), select_point=["tail", "ear"])
    assert c2["index_list"] == [1, 4]
    c3 = F.select_points(c, ["paw_l"])
    assert c3["index_list"] == [2] and c["index_list"] is None and c3["pred_tracks"] is not c["pred_tracks"]
    assert F.select_points(c2, None)["index_list"] is None
    with pytest.raises(KeyError):
        F.load_tap(p, select_point=["wing"])
    with pytest.raises(KeyError, match="point_embedding"):
        F.load_tap({"pred_tracks": tracks, "point_name2id": {}})
    with pytest.raises(ValueError, match="5 points"):
        F.load_tap({"pred_tracks": tracks, "point_name2id": {}, "point_embedding": torch.zeros(4, 1280)})


def test_motion_module_checkpoint_loads_into_the_unet(tmp_path):
    """test.py:60-64: an AnimateDiff checkpoint (keys with `.pos_encoder.pe`) loads after the key remap; every motion weight
    -- including the zero-initialised proj_out -- takes the checkpoint's value."""
    from videoswap_b200 import AnimateDiffUNet3DModel
    unet = AnimateDiffUNet3DModel.from_config(dict(block_out_channels=(32, 64, 128, 128), cross_attention_dim=64, norm_num_groups=8,
                                                   attention_head_dim=4, use_motion_module=True, motion_module_type="Vanilla",
                                                   motion_module_kwargs=dict(num_attention_heads=4, temporal_position_encoding_max_len=24)))
    sd = unet.state_dict()
    motion = {k: v for k, v in sd.items() if "motion_modules" in k}
    assert motion and any(k.endswith(".processor.pos_encoder.pe") for k in motion)
    po = next(k for k in motion if k.endswith("temporal_transformer.proj_out.weight"))
    assert float(sd[po].abs().max()) == 0.0                                         # reference initial state
    g = torch.Generator().manual_seed(5)
    ckpt = {k.replace(".processor.pos_encoder", ".pos_encoder"): torch.randn(v.shape, generator=g).to(v.dtype) for k, v in motion.items()}
    assert not any(".processor." in k for k in ckpt)
    p = tmp_path / "mm_sd_v15_v2.ckpt"
    torch.save(ckpt, p)
    r = F.load_motion_module(unet, p)
    missing = list(r.missing_keys if hasattr(r, "missing_keys") else r[0])
    assert missing and not any("motion_modules" in k for k in missing)              # only the 2-D weights are "missing"
    after = unet.state_dict()
    for k, v in ckpt.items():
        assert torch.equal(after[k.replace(".pos_encoder", ".processor.pos_encoder")], v), k
    with pytest.raises(KeyError, match="does not have"):
        F.load_motion_module(unet, {"down_blocks.0.motion_modules.9.nothing.weight": torch.zeros(1)})


def test_adapter_checkpoint_roundtrip(tmp_path):
    from videoswap_b200 import SparsePointAdapter
    a = SparsePointAdapter()
    sd = {k: torch.randn_like(v) for k, v in a.state_dict().items()}
    p = tmp_path / "adapter.pth"
    torch.save(sd, p)
    b = F.load_adapter(SparsePointAdapter(), p, dtype=torch.float16)
    for k, v in b.state_dict().items():
        assert v.dtype == torch.float16 and torch.equal(v, sd[k].half())
    with pytest.raises(RuntimeError):
        F.load_adapter(SparsePointAdapter(), {"nope": torch.zeros(1)})
