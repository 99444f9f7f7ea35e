"""MFCC / class constants of the reference (config.py:19-27,45-47); the dataset paths of the
reference's config.py are the author's laptop and are not reproduced."""
SAMPLERATE = 16000
FRAME_SIZE = 400
FRAME_STEP = 160
LOW_HZ = 300
HIGH_HZ = 8000
FILTERBANKS_NUM = 26
MFCC_NUM = 13
FFT_N = 512

PROCESSES_NUM = 4      # config.py:30 (reference Pool size; the GPU path shards by utterance instead)
FILES_PER_STEP = 30    # config.py:32

NONE_VOICED = 0
VOICED = 1
MUSIC = 2
