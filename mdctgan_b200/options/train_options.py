"""TrainOptions (reference: options/train_options.py) -- also what generate_audio.py parses (isTrain=True,
generate_audio.py:13)."""
from .base_options import BASE_FLAGS, EXTRA_FLAGS, TRAIN_FLAGS, BaseOptions


class TrainOptions(BaseOptions):
    isTrain = True

    def flag_table(self):
        return BASE_FLAGS + TRAIN_FLAGS + EXTRA_FLAGS
