"""Drives the UNMODIFIED reference (neoncloud/mdctGAN) for bench.py's reference legs -- BASELINE INFRASTRUCTURE ONLY.

`install()` copies the reference's Python packages (models/, util/, options/, data/) from /root/reference into the
git-ignored `baseline/_ref/` (the reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref`
has nothing to build; a plain copy is the equivalent install).  `baseline/_ref` is NOT gpurun-ignored, so it travels to
the GPU box, where /root/reference does not exist.  Nothing from it is tracked by git.

`make_stepper(device)` builds the reference's own `create_model(TrainOptions().parse())` (Pix2PixHDModel with its
Audio2MDCT, LocalEnhancer, MultiscaleDiscriminator, GANLoss, two torch.optim.Adam) and returns a closure that runs one
iteration of /root/reference/train.py:160-202 verbatim (model._forward -> loss_G.backward() -> optimizer_G.step() ->
loss_D.backward() -> optimizer_D.step()).  On "cpu" this is the reference arm of the bench (kind "reference"); on
"cuda" it is the torch-eager cuDNN / cuFFT path of the same code -- the GPU "kernel to beat" (SURVEY.md 8d).

Three modules the reference imports are absent from this image: torch_scatter and matplotlib (never
called on this path -> empty stubs) and bottleneck_transformer_pytorch==0.1.4 (-> oracle/bottlestack_ref.py, our
restatement of the package).  No kernel, model or engine of mdctgan_b200 is on this path.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("MDCTGAN_REFERENCE", "/root/reference")
PACKAGES = ("models", "util", "options", "data")


def install(force=False):
    """Copy the reference packages into baseline/_ref (only where /root/reference exists: the build container)."""
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(os.path.join(REF_DST, "models"))
    if force and os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    for pkg in PACKAGES:
        dst = os.path.join(REF_DST, pkg)
        if not os.path.isdir(dst):
            shutil.copytree(os.path.join(REF_SRC, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.csv", "*.sh"))
    return True


def available():
    return os.path.isdir(os.path.join(REF_DST, "models"))


_imported = False


def import_reference():
    global _imported
    if _imported:
        return
    if not available():
        raise RuntimeError("baseline/_ref is missing (run __graft_entry__.build() where /root/reference exists)")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    ts = types.ModuleType("torch_scatter")
    ts.scatter = None
    sys.modules.setdefault("torch_scatter", ts)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:  # noqa: BLE001
            mp, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            pp.switch_backend = lambda *a, **k: None
            mp.pyplot = pp
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mp, pp
    try:
        import bottleneck_transformer_pytorch  # noqa: F401
    except Exception:  # noqa: BLE001
        import oracle.bottlestack_ref as bs

        sys.modules["bottleneck_transformer_pytorch"] = bs
    # the reference's top-level package names (models, util, options, data) must win over anything else on sys.path
    for name in list(sys.modules):
        if name.split(".")[0] in PACKAGES and not getattr(sys.modules[name], "__file__", "").startswith(REF_DST):
            del sys.modules[name]
    sys.path.insert(0, REF_DST)
    _imported = True


def ref_options(extra_args, device):
    """TrainOptions().parse() of the reference under a patched sys.argv (it writes opt.txt under --checkpoints_dir)."""
    import_reference()
    from options.train_options import TrainOptions

    tmp = tempfile.mkdtemp(prefix="mdctgan_ref_")
    gpu = "-1" if str(device) == "cpu" else str(getattr(device, "index", 0) or 0)
    argv = ["x", "--checkpoints_dir", tmp, "--gpu_ids", gpu] + list(extra_args)
    old = sys.argv
    sys.argv = argv
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            opt = TrainOptions().parse()
    finally:
        sys.argv = old
    return opt


def make_stepper(opt_args, lr_audio, hr_audio, device="cpu", seed=1234, fp16=False):
    """Returns (step, model): step() runs one train.py:160-202 iteration on the given batch and returns the four losses
    [G_GAN, G_GAN_Feat, D_real, D_fake] as floats (the reference's loop reads them every print_freq steps; here every step,
    like the GPU arm's e2e leg)."""
    import torch

    opt = ref_options(opt_args, device)
    from models.models import create_model

    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(opt)
    model.train()
    dev = torch.device(device)
    lr_t, hr_t = torch.as_tensor(lr_audio).to(dev), torch.as_tensor(hr_audio).to(dev)
    optimizer_G, optimizer_D = model.optimizer_G, model.optimizer_D
    scaler = torch.amp.GradScaler("cuda") if fp16 else None

    def step(read_losses=True):
        if fp16:
            with torch.autocast("cuda"):
                losses, _ = model._forward(lr_t, hr_t, infer=False)
        else:
            losses, _ = model._forward(lr_t, hr_t, infer=False)
        losses = [torch.mean(x) if not isinstance(x, int) else x for x in losses]
        loss_dict = dict(zip(model.loss_names, losses))
        loss_D = (loss_dict["D_fake"] + loss_dict["D_real"]) * 0.5
        loss_G = loss_dict["G_GAN"] + loss_dict.get("G_GAN_Feat", 0)
        optimizer_G.zero_grad()
        if fp16:
            scaler.scale(loss_G).backward()
            scaler.step(optimizer_G)
        else:
            loss_G.backward()
            optimizer_G.step()
        optimizer_D.zero_grad()
        if fp16:
            scaler.scale(loss_D).backward()
            scaler.step(optimizer_D)
            scaler.update()
        else:
            loss_D.backward()
            optimizer_D.step()
        if read_losses:
            return [float(loss_dict[k].detach()) for k in ("G_GAN", "G_GAN_Feat", "D_real", "D_fake")]
        return None

    return step, model
