"""Decoding environment with the reference's Decoder API (reference: neuralNetworks/decoder.py).

`Decoder(classifier, input_dim, max_length)(inputs)` returned softmax posteriors from
`outputs.eval(feed)` on the utterance padded to max_length (decoder.py:49-71); here it is one
tfk_forward_posteriors call on the unpadded frames.  `loglik` additionally fuses Nnet.decode's
host-side `np.log(output / prior)` (nnet.py:280-286) into the output kernel."""
import numpy as np

from ..engine import Engine


class Decoder(object):
    def __init__(self, classifier, input_dim, max_length, *, precision="bf16x3", device=None, max_frames=16384):
        """precision defaults to the fp32-equivalent mode: posteriors / log-likelihoods are what the 1e-3 contract
        against the reference's fp32 TensorFlow arithmetic is stated on (decoder.py:26-27, 44); "bf16" is 3x faster
        and measured at full size in tests/test_gpu_zfullsize.py (log-likelihood error ~1e-2 absolute)."""
        spec = classifier.engine_spec(input_dim)
        self.max_length = max_length
        self.input_dim = input_dim
        self.engine = Engine(spec["num_layers"], spec["input_dim"], spec["hidden_dim"], spec["output_dim"],
                             min(int(max_frames), max(int(max_length), 1)), nonlin=spec["nonlin"],
                             batch_norm=spec["batch_norm"], keep_prob=spec["keep_prob"], l2_norm=spec["l2_norm"], precision=precision, device=device)
        self._trainer_like = None

    def __call__(self, inputs):
        """[N, F] numpy -> [N, O] numpy posteriors (eval mode: moving-stat batch norm, no dropout)"""
        return self.engine.posteriors(np.ascontiguousarray(inputs, dtype=np.float32)).cpu().numpy()

    def loglik(self, inputs, prior, out=None):
        """device tensor [N, O] = log(softmax / prior), no flooring (nnet.py:280-286)"""
        return self.engine.loglik(inputs, prior, out=out)

    def restore(self, filename):
        """load the model written by Trainer.save_model — or by the reference's own saver (decoder.py:73-81)"""
        from .trainer import MODEL_NAMES, read_model_file

        params = {}
        for key, val in read_model_file(filename).items():
            if key == "Classifier/initialisedlayers":
                self.engine.set_active_layers(int(val) + 1)
                continue
            if not key.startswith("Classifier/"):
                continue
            _, layer, rest = key.split("/", 2)
            params[MODEL_NAMES[rest] + layer[len("layer"):]] = val
        self.engine.load_params(params)
