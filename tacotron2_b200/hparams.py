"""TF-free stand-in for the reference ``hparams.py`` (which needs tensorflow 1.x's
``tf.contrib.training.HParams``): same attribute names and defaults (hparams.py:12-85) and the same
``--hparams=a=b,c=d`` override syntax (hparams.py:88-90)."""
from types import SimpleNamespace

N_SYMBOLS = 148   # len(text.symbols): pad + '-' + 10 punctuation + 52 letters + 84 ARPAbet (text/symbols.py:9-18)


class HParams(SimpleNamespace):
    def values(self):
        return dict(vars(self))

    def parse(self, string):
        """'k=v,k2=v2' overrides; values are cast to the type of the existing default."""
        if not string:
            return self
        for item in string.split(","):
            if not item.strip():
                continue
            k, v = item.split("=", 1)
            k = k.strip()
            if not hasattr(self, k):
                raise ValueError("unknown hparam %r" % k)
            cur = getattr(self, k)
            if isinstance(cur, bool):
                v = v.strip().lower() in ("1", "true", "yes")
            elif isinstance(cur, int):
                v = int(v)
            elif isinstance(cur, float):
                v = float(v)
            elif isinstance(cur, list):
                v = [x for x in v.strip("[]").split(";") if x]
            setattr(self, k, v)
        return self


def create_hparams(hparams_string=None, verbose=False):
    hp = HParams(
        # experiment (hparams.py:12-22)
        epochs=500, iters_per_checkpoint=1000, seed=1234, dynamic_loss_scaling=True,
        fp16_run=False, distributed_run=False, dist_backend="nccl",
        dist_url="tcp://localhost:54321", cudnn_enabled=True, cudnn_benchmark=False,
        ignore_layers=['embedding.weight'],
        # data (hparams.py:27-30)
        load_mel_from_disk=False,
        training_files='filelists/ljs_audio_text_train_filelist.txt',
        validation_files='filelists/ljs_audio_text_val_filelist.txt',
        text_cleaners=['english_cleaners'],
        # audio (hparams.py:35-42)
        max_wav_value=32768.0, sampling_rate=22050, filter_length=1024, hop_length=256,
        win_length=1024, n_mel_channels=80, mel_fmin=0.0, mel_fmax=8000.0,
        # model (hparams.py:47-75)
        n_symbols=N_SYMBOLS, symbols_embedding_dim=512,
        encoder_kernel_size=5, encoder_n_convolutions=3, encoder_embedding_dim=512,
        n_frames_per_step=1, decoder_rnn_dim=1024, prenet_dim=256, max_decoder_steps=1000,
        gate_threshold=0.5, p_attention_dropout=0.1, p_decoder_dropout=0.1,
        attention_rnn_dim=1024, attention_dim=128,
        attention_location_n_filters=32, attention_location_kernel_size=31,
        postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5,
        # optimisation (hparams.py:80-85)
        use_saved_learning_rate=False, learning_rate=1e-3, weight_decay=1e-6,
        grad_clip_thresh=1.0, batch_size=64, mask_padding=True)
    if hparams_string:
        hp.parse(hparams_string)
    if verbose:
        print("Final parsed hparams:", hp.values())
    return hp
