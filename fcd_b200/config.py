"""Default parameter dict: the keys of the reference's config.py:1-69 that the hot path reads, same defaults."""


def get_default_params():
    return {
        "model_type": "MS_DSA_NET", "model_returns_vaeloss": False, "sa_type": "parallel", "feature_size": 16,
        "project_size": 64, "patch_size": 128, "chans_in": 2, "chans_out": 2, "batch_size": 1, "use_amp": True,
        "min_region_size": 50, "seed": 42, "lr": 1e-4, "weight_decay": 1e-5, "loss": "DiceLoss", "lambda_dice": 1.0,
        "lambda_ce": 1.0, "lambda_focal": 1.0, "ce_background_weight": 0.5, "ce_fcd_weight": 0.5,
        "gamma_focal": 2.0, "gdice_wtype": "square", "jaccard": False, "square_pred": False, "sigmoid": False,
        "softmax": True, "segresnet_upsample_mode": "pixelshuffle", "segresnet_deeper": False,
        "tv_loss_norm": "l1", "tv_loss_weight": 0.0, "tvloss_exclude_borders": False, "boundaryloss_weight": 0.0,
        "loss_vae_weight": 0.2,
    }
