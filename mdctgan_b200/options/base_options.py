"""The reference's command-line surface (options/base_options.py:11-127, options/train_options.py:5-73) as
one flag table: same names, types and defaults, so scripts written for train.py / generate_audio.py parse
unchanged.  `parse()` returns the Namespace that `models.create_model(opt)` consumes."""
import argparse
import os

import torch

from .audio_config import BINS, FRAME_LENGTH, HOP_LENGTH, HR_SAMPLE_RATE, LR_SAMPLE_RATE, N_FFT, SR_SAMPLE_RATE, WIN_LENGTH

FLAG = "flag"   # action='store_true'


def _key_map(text):
    return {str(k): str(v) for k, v in (item.split(":") for item in text.split(","))}


# (name, type | FLAG, default[, extra argparse kwargs])
BASE_FLAGS = [
    ("name", str, "label2city"), ("gpu_ids", str, "0"), ("checkpoints_dir", str, "./checkpoints"), ("model", str, "pix2pixHD"),
    ("norm", str, "instance"), ("use_dropout", FLAG, False), ("data_type", int, 32, dict(choices=[8, 16, 32])), ("verbose", FLAG, False),
    ("fp16", FLAG, False), ("local_rank", int, 0), ("seed", int, 42), ("fit_residual", FLAG, False),
    ("batchSize", int, 1), ("loadSize", int, 1024), ("fineSize", int, 512), ("label_nc", int, 0), ("input_nc", int, 2), ("output_nc", int, 1),
    ("dataroot", str, "./datasets/vctk/train.csv"), ("evalroot", str, "./datasets/vctk/test.csv"), ("serial_batches", FLAG, False),
    ("nThreads", int, 2), ("max_dataset_size", int, float("inf")),
    ("explicit_encoding", FLAG, False), ("alpha", float, 0.6), ("norm_range", float, (0, 1), dict(nargs=2)), ("abs_norm", FLAG, False),
    ("src_range", float, (-5, 5), dict(nargs=2)), ("arcsinh_transform", FLAG, False), ("raw_mdct", FLAG, False), ("arcsinh_gain", float, 500),
    ("add_noise", FLAG, False), ("snr", float, 55), ("display_winsize", int, 512), ("tf_log", FLAG, False),
    ("netG", str, "global"), ("ngf", int, 64), ("upsample_type", str, "transconv"), ("downsample_type", str, "conv"),
    ("n_downsample_global", int, 4), ("n_blocks_global", int, 9), ("n_blocks_attn_g", int, 1), ("proj_factor_g", int, 4),
    ("dim_head_g", int, 128), ("heads_g", int, 4), ("n_blocks_local", int, 3), ("n_blocks_attn_l", int, 0), ("proj_factor_l", int, 4),
    ("dim_head_l", int, 128), ("heads_l", int, 4), ("n_local_enhancers", int, 1), ("niter_fix_global", int, 0),
    ("mask", FLAG, False), ("smooth", float, 0.0), ("mask_hr", FLAG, False), ("mask_mode", str, None), ("min_value", float, 1e-7),
]

TRAIN_FLAGS = [
    ("display_freq", int, 200), ("print_freq", int, 100), ("save_latest_freq", int, 1000), ("save_epoch_freq", int, 10),
    ("eval_freq", int, 32000), ("loss_update_freq", int, 256), ("no_html", FLAG, False), ("debug", FLAG, False), ("abs_spectro", FLAG, False),
    ("continue_train", FLAG, False), ("freeze_g_d", FLAG, False), ("freeze_g_u", FLAG, False), ("freeze_l_d", FLAG, False),
    ("freeze_l_u", FLAG, False), ("load_pretrain", str, ""), ("param_key_map", _key_map, {}), ("which_epoch", str, "latest"),
    ("phase", str, "train"), ("niter", int, 100), ("niter_decay", int, 100), ("niter_limit_aux", int, 20), ("beta1", float, 0.5),
    ("lr", float, 0.0002), ("validation_split", float, 0.05), ("val_indices", str, None), ("eval_size", int, 100),
    ("phase_encoding_mode", str, None), ("num_D", int, 2), ("n_layers_D", int, 3), ("ndf", int, 64), ("no_ganFeat_loss", FLAG, False),
    ("lambda_feat", float, 10.0), ("no_lsgan", FLAG, False), ("pool_size", int, 0),
    ("lr_sampling_rate", int, LR_SAMPLE_RATE), ("hr_sampling_rate", int, HR_SAMPLE_RATE), ("sr_sampling_rate", int, SR_SAMPLE_RATE),
    ("segment_length", int, FRAME_LENGTH), ("gen_overlap", int, 0), ("n_fft", int, N_FFT), ("bins", int, BINS), ("hop_length", int, HOP_LENGTH),
    ("win_length", int, WIN_LENGTH), ("center", FLAG, False), ("is_lr_input", FLAG, False),
]

# ours, not in the reference: arithmetic flavour of the fused transform kernels
EXTRA_FLAGS = [("mdct_precision", str, "fp32", dict(choices=["fp32", "fp64"]))]


def _add(parser, spec):
    name, kind, default = spec[:3]
    extra = spec[3] if len(spec) > 3 else {}
    if kind is FLAG:
        parser.add_argument("--" + name, action="store_true", default=default)
    else:
        parser.add_argument("--" + name, type=kind, default=default, **extra)


class BaseOptions:
    isTrain = False

    def __init__(self):
        self.parser = argparse.ArgumentParser()
        self.initialized = False

    def flag_table(self):
        return BASE_FLAGS + EXTRA_FLAGS

    def initialize(self):
        for spec in self.flag_table():
            _add(self.parser, spec)
        self.initialized = True

    def parse(self, save=True, args=None):
        if not self.initialized:
            self.initialize()
        opt = self.parser.parse_args(args)
        opt.isTrain = self.isTrain
        opt.gpu_ids = [int(s) for s in opt.gpu_ids.split(",") if int(s) >= 0]
        if opt.gpu_ids and torch.cuda.is_available():
            torch.cuda.set_device(opt.gpu_ids[0])
        if save and not getattr(opt, "continue_train", False):
            expr_dir = os.path.join(opt.checkpoints_dir, opt.name)
            os.makedirs(expr_dir, exist_ok=True)
            with open(os.path.join(expr_dir, "opt.txt"), "wt") as fh:
                fh.write("------------ Options -------------\n")
                for k, v in sorted(vars(opt).items()):
                    fh.write("%s: %s\n" % (str(k), str(v)))
                fh.write("-------------- End ----------------\n")
        self.opt = opt
        return opt
