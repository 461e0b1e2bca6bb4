"""Serving path: ``GlowTTS.inference`` as one CUDA graph.

The reference's ``inference`` (Modules.py:128-204) reads the predicted mel lengths back to the host to size its
tensors, so every call is a chain of small launches with host round trips in between.  ``GraphedInference``
captures ``GlowTTS.inference_device`` -- the same arithmetic with static shapes (B, T_text_max, T_mel_max), the
lengths kept on the device and the decoder on a fixed-geometry row map -- once, and replays it per request
batch: one pinned host->device copy of the tokens, one graph launch, and the caller reads back what it needs.
"""
import torch

from . import _lib


class GraphedInference:
    def __init__(self, model, batch, max_text_length, max_mel_length=1000, speakers=False, noise_scale=1.0,
                 length_scale=1.0, warmup=2, device=None):
        dev = torch.device(device if device is not None else next(model.parameters()).device)
        if dev.type != "cuda":
            raise _lib.GlowCoreError("GraphedInference needs the model on a CUDA device")
        self.model, self.device = model, dev
        self.batch, self.t_text, self.t_mel = int(batch), int(max_text_length), int(max_mel_length)
        self.noise_scale = float(noise_scale)
        self.tokens = torch.ones((self.batch, self.t_text), dtype=torch.int64, device=dev)      # 1 = <E>, the pad token
        self.lengths = torch.full((self.batch,), 2, dtype=torch.int32, device=dev)
        self.speakers = torch.zeros((self.batch,), dtype=torch.int64, device=dev) if speakers else None
        self.length_scale = torch.full((self.batch,), float(length_scale), dtype=torch.float32, device=dev)
        self._pin_tokens = torch.ones((self.batch, self.t_text), dtype=torch.int64).pin_memory()
        self._pin_lengths = torch.full((self.batch,), 2, dtype=torch.int32).pin_memory()
        model.eval()
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        # the graph holds raw pointers into cache-owned objects (workspace, device row map): keep them alive with it
        with _lib.capture_keepalive() as keep:
            with torch.cuda.stream(side):                  # eager warm-up: weight packs, workspaces, library handles
                for _ in range(max(1, warmup)):
                    self._call()
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.mels, self.mel_lengths, self.attentions = self._call()
        self._keep = list(keep)
        self.launches_per_replay = _lib.launch_count() - before

    def _call(self):
        return self.model.inference_device(self.tokens, self.lengths, speakers=self.speakers,
                                           noise_scale=self.noise_scale, length_scale=self.length_scale,
                                           max_mel_length=self.t_mel)

    def load(self, tokens, token_lengths, speakers=None):
        """tokens [b <= batch, t <= max_text_length] int64 and lengths [b] on the host (unused slots keep a
        2-token dummy sentence); stream-ordered pinned copies."""
        b, t = tokens.shape
        if b > self.batch or t > self.t_text:
            raise ValueError("request (%d x %d) exceeds the captured shape (%d x %d)" % (b, t, self.batch, self.t_text))
        self._pin_tokens.fill_(1)
        self._pin_tokens[:b, :t] = tokens
        self._pin_lengths.fill_(2)
        self._pin_lengths[:b] = token_lengths.to(torch.int32)
        self.tokens.copy_(self._pin_tokens, non_blocking=True)
        self.lengths.copy_(self._pin_lengths, non_blocking=True)
        if self.speakers is not None and speakers is not None:
            self.speakers[:b].copy_(speakers.to(torch.int64), non_blocking=True)

    def run(self, tokens=None, token_lengths=None, speakers=None):
        """-> (mels [batch,80,max_mel_length], mel_lengths [batch] int64), device tensors owned by the graph
        (overwritten by the next run)."""
        if tokens is not None:
            self.load(tokens, token_lengths, speakers)
        self.graph.replay()
        return self.mels, self.mel_lengths
