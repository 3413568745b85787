"""Model hyper-parameters of the shipped ViewCrafter configs (reference: configs/inference_pvd_1024.yaml:36-88,
configs/inference_pvd_512.yaml).  The 512 and 1024 models share one architecture; only image_size / base_scale differ."""

UNET_PARAMS = dict(in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                   num_res_blocks=2, channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64,
                   transformer_depth=1, context_dim=1024, use_linear=True, use_checkpoint=False,
                   temporal_conv=True, temporal_attention=True, temporal_selfatt_only=True,
                   use_relative_position=False, use_causal_attention=False, temporal_length=16,
                   addition_attention=True, image_cross_attention=True, default_fs=10, fs_condition=True)

VAE_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                    ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)

DIFFUSION = {
    "ViewCrafter_25": dict(image_size=(72, 128), frames=25, base_scale=0.3),        # 576x1024
    "ViewCrafter_25_512": dict(image_size=(40, 64), frames=25, base_scale=0.7),     # 320x512
    "ViewCrafter_16": dict(image_size=(72, 128), frames=16, base_scale=0.3),
}
SCHEDULE = dict(timesteps=1000, linear_start=0.00085, linear_end=0.012, rescale_betas_zero_snr=True,
                parameterization="v", scale_factor=0.18215, turning_step=400)
