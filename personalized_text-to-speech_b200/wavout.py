"""Output side of the path: what cmd_inference.py:114-117 / VC_inference.py:49-51 do with the
decoder's result -- ``[0][0,0].data.cpu().float().numpy()`` then ``scipy.io.wavfile.write(path, 22050, audio)`` --
for batches, without the synchronous ``.cpu()`` and without a numpy detour.

``WavBatchWriter`` keeps one PINNED host buffer per in-flight utterance laid out as a complete RIFF/WAVE file: the
header is written once on the host, the samples arrive by an asynchronous D2H copy straight behind it (float32: the
decoder's output as it is, the format the reference writes; pcm16: converted on the device by ``vitsdec_wav_pcm16`` so
the copy moves half the bytes), and saving is one ``file.write`` of the buffer.  The bytes are identical to what
``scipy.io.wavfile.write`` produces for the same samples (tests/test_generator_host.py, tests/test_gpu_decoder.py).
"""
import struct

import numpy as np
import torch

from . import _capi


def wav_header(n_samples, sample_rate=22050, fmt="float32"):
    """RIFF/WAVE header for a mono file, byte for byte what scipy.io.wavfile.write emits: PCM = 44 bytes; IEEE float
    adds the cbSize field and a 'fact' chunk (58 bytes)."""
    if fmt == "float32":
        tag, bits = 3, 32
    elif fmt == "pcm16":
        tag, bits = 1, 16
    else:
        raise ValueError("fmt must be 'float32' or 'pcm16'")
    nbytes = n_samples * (bits // 8)
    fmt_chunk = struct.pack("<HHIIHH", tag, 1, sample_rate, sample_rate * (bits // 8), bits // 8, bits)
    if tag != 1:
        fmt_chunk += b"\x00\x00"
    h = b"RIFF" + b"\x00\x00\x00\x00" + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt_chunk)) + fmt_chunk
    if tag != 1:
        h += b"fact" + struct.pack("<II", 4, n_samples)
    h += b"data" + struct.pack("<I", min(nbytes, 0xFFFFFFFF))
    total = len(h) + nbytes
    return h[:4] + struct.pack("<I", total - 8) + h[8:]


def pcm16_reference(x):
    """The conversion vitsdec_wav_pcm16 performs, in numpy (tests): round-to-nearest-even(clip(x, -1, 1) * 32767)."""
    return np.rint(np.clip(np.asarray(x, dtype=np.float32), -1.0, 1.0) * np.float32(32767.0)).astype(np.int16)


class WavBatchWriter:
    """Turn decoded waveforms [B, 1, L] (device, fp32) into WAV file images in pinned host memory, asynchronously."""

    def __init__(self, sample_rate=22050, fmt="float32"):
        self.sample_rate = int(sample_rate)
        self.fmt = fmt
        self.hdr = len(wav_header(0, sample_rate, fmt))
        self.itemsize = 4 if fmt == "float32" else 2
        self._pool = {}   # (B, L) -> list of free pinned buffers

    def _buffer(self, n):
        # one pinned byte buffer per utterance, padded in front so that the SAMPLE region is 16-byte aligned (the header
        # is 44 / 58 bytes): file image = buf[pad : pad + hdr + n * itemsize]
        pad = (-self.hdr) % 16
        buf = torch.empty(pad + self.hdr + n * self.itemsize, dtype=torch.uint8).pin_memory()
        buf[pad:pad + self.hdr] = torch.frombuffer(bytearray(wav_header(n, self.sample_rate, self.fmt)), dtype=torch.uint8)
        return buf, pad

    def enqueue(self, y, lengths=None, stream=None):
        """y: [B, 1, L] fp32 on a CUDA device.  lengths: valid samples per utterance (None = L).  Issues the conversion
        and the D2H copies on `stream` (default: current) and returns (images, event): images[i] is a uint8 CPU tensor
        holding the complete file of utterance i once `event` has completed."""
        if not (y.is_cuda and y.dtype == torch.float32 and y.dim() == 3 and y.shape[1] == 1):
            raise RuntimeError("WavBatchWriter.enqueue: expected a CUDA fp32 tensor [B, 1, L]")
        B, _, L = y.shape
        stream = stream or torch.cuda.current_stream(y.device)
        y = y.contiguous()
        images = []
        with torch.cuda.stream(stream):
            src = y
            if self.fmt == "pcm16":
                src = torch.empty((B, 1, L), dtype=torch.int16, device=y.device)
                lib = _capi.lib()
                _capi.check(lib.vitsdec_wav_pcm16(y.device.index or 0, y.data_ptr(), src.data_ptr(), B * L,
                                                  stream.cuda_stream), "vitsdec_wav_pcm16")
                src.record_stream(stream)
            for i in range(B):
                n = L if lengths is None else int(lengths[i])
                buf, pad = self._buffer(n)
                dst = buf[pad + self.hdr:].view(torch.float32 if self.fmt == "float32" else torch.int16)
                dst.copy_(src[i, 0, :n], non_blocking=True)
                images.append(buf[pad:])
            ev = torch.cuda.Event()
            ev.record(stream)
        y.record_stream(stream)
        return images, ev

    @staticmethod
    def save(images, paths, event=None):
        if event is not None:
            event.synchronize()
        for img, path in zip(images, paths):
            with open(path, "wb") as f:
                f.write(img.numpy().tobytes())


def synthesize_to_wav(net_g, x, x_lengths, sid, paths, sample_rate=22050, fmt="float32", **infer_kwargs):
    """Batched equivalent of cmd_inference.py:108-117: ``net_g`` is the reference's SynthesizerTrn (with ``dec``
    patched to the B200 Generator or not), x / x_lengths / sid a BATCH of symbol-id sequences; every utterance is
    trimmed to its own length (``y_mask``) and written to paths[i].  Returns the per-utterance sample counts."""
    hop = getattr(net_g.dec, "hop", 256)
    with torch.no_grad():
        o, _, y_mask, _ = net_g.infer(x, x_lengths, sid=sid, **infer_kwargs)
    lengths = (y_mask.sum(dim=(1, 2)).long() * hop).tolist()
    w = WavBatchWriter(sample_rate, fmt)
    images, ev = w.enqueue(o.float(), lengths)
    w.save(images, paths, ev)
    return lengths
