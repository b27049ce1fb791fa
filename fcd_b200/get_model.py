"""get_model(params) with the reference's signature and side effects (get_model.py:9-249), building fcd_b200
modules whose state-dict keys / shapes equal the reference's so reference checkpoints load unchanged."""
from __future__ import annotations

_ACT = ("leakyrelu", {"inplace": True, "negative_slope": 0.01})


def _blocks(params):
    deeper = params.get("segresnet_deeper", False)
    return ((1, 2, 2, 4, 4), (2, 2, 2, 2)) if deeper else ((1, 2, 2, 4), (1, 1, 1))


def get_model(params, return_model=True):
    from . import networks as N
    model = None
    params["model_returns_vaeloss"] = False
    mt = params["model_type"].lower()
    if mt in ("ms_dsa_net", "ms_dsa_net_ps"):
        if return_model:
            kw = dict(spatial_dims=3, in_channels=params["chans_in"], out_channels=params["chans_out"],
                      img_size=params["patch_size"], feature_size=params["feature_size"], pos_embed=True,
                      project_size=params["project_size"], sa_type=params["sa_type"], norm_name="instance",
                      act_name=_ACT, res_block=True, bias=False, dropout_rate=0.1)
            if mt == "ms_dsa_net":
                model = N.MS_DSA_NET(**kw)
            else:
                model = N.MS_DSA_NET_PS(**kw, upsample_mode="pixelshuffle", interpolate_mode="linear")
    elif mt == "baseunet":
        if return_model:
            model = N.BaseUNet(spatial_dims=3, in_channels=params["chans_in"], out_channels=params["chans_out"],
                               feature_size=params["feature_size"], norm_name="instance", act_name=_ACT,
                               res_block=True, bias=False, depth=6)
    elif mt in ("segresnet", "segresnetvae", "segresnet_dsa", "segresnetvae_dsa"):
        blocks_down, blocks_up = _blocks(params)
        common = dict(spatial_dims=3, in_channels=params["chans_in"], out_channels=params["chans_out"],
                      init_filters=params["feature_size"], dropout_prob=0.1, norm="INSTANCE", use_conv_final=True,
                      upsample_mode=params["segresnet_upsample_mode"], blocks_down=blocks_down, blocks_up=blocks_up)
        vae = dict(input_image_size=params["patch_size"], vae_estimate_std=False, vae_default_std=0.3, vae_nz=256)
        dsa = dict(dsa_img_size=params["patch_size"], dsa_project_size=params["project_size"], dsa_num_heads=4,
                   dsa_pos_embed=True, dsa_dropout_rate=0.1, dsa_sa_type=params["sa_type"], dsa_bias=False,
                   dsa_num_layers=3, dsa_start_level=len(blocks_down) - 2)
        if return_model:
            if mt == "segresnet":
                model = N.SegResNet(act=("RELU", {"inplace": True}), **common)
            elif mt == "segresnetvae":
                model = N.SegResNetVAE(**vae, **common)
            elif mt == "segresnet_dsa":
                model = N.SegResNet_DSA(**common, **dsa)
            else:
                model = N.SegResNetVAE_DSA(**vae, **common, **dsa)
        if mt in ("segresnetvae", "segresnetvae_dsa"):
            params["model_returns_vaeloss"] = True
    else:
        raise NotImplementedError(
            f"model_type {params['model_type']!r} is outside the fcd_b200 hot-path scope (SURVEY.md section 2); "
            "in scope: ms_dsa_net, ms_dsa_net_ps, baseunet, segresnet, segresnetvae, segresnet_dsa, segresnetvae_dsa")
    if model is not None:
        n = sum(p.numel() for p in model.parameters() if p.requires_grad)
        print(f"Trainable parameters: {n}")
    return model, params
