"""create_model(opt) with the reference's contract (models/models.py:3-20)."""


def create_model(opt):
    if opt.model != "pix2pixHD":
        raise NotImplementedError("model [%s]: only 'pix2pixHD' exists on the audio path (the reference's UIModel is image-only)" % opt.model)
    from .pix2pixHD_model import InferenceModel, Pix2PixHDModel

    model = Pix2PixHDModel() if opt.isTrain else InferenceModel()
    model.initialize(opt)
    if getattr(opt, "verbose", False):
        print("model [%s] was created" % (model.name()))
    return model
