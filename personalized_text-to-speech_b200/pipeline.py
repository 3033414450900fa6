"""Host-buffer decode pipeline: overlaps the H2D copy of the next batch and the D2H copy of the previous waveform with
the decode of the current one (the output side of cmd_inference.py:110-114).

The reference moves one batch at a time: ``.to(device)`` ... ``.data.cpu().float().numpy()``.  With the decoder at
~10 ms per 16 x 10 s batch, those copies (10.6 MB in, 14.1 MB out over PCIe) are 4 % of the step when they are
serialised with it.  ``HostPipeline`` keeps ``depth`` batches in flight on separate CUDA streams (the copy engines run
beside the SMs); inputs and outputs are PINNED host tensors, results are complete after ``wait`` / ``wait_all``.
"""
import torch


class HostPipeline:
    def __init__(self, generator, depth=2, device=None):
        self.G = generator
        p = next(generator.parameters())
        self.device = torch.device(device) if device is not None else p.device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline: the generator must live on a CUDA device (there is no CPU path)")
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, int(depth)))]
        self.done = [None] * len(self.streams)
        self._n = 0

    def submit(self, z_host, g_host, out_host, on_device=None, decode_after=None):
        """Enqueue H2D(z, g) -> decode -> D2H(out) on the next stream.  z_host [B, C, T], g_host [B, gin, 1] or None and
        out_host [B, 1, T*hop] are pinned CPU tensors; out_host is valid after wait(ticket).  Returns a ticket.

        on_device(y, stream): optional hook called with the decoded waveform still on the device, inside the slot's
        stream context right after the decode was enqueued -- e.g. to start the multi-GPU waveform gather on a side
        stream (bench.py) or a WavBatchWriter.enqueue -- so that it overlaps the D2H copy and the next batch.

        decode_after: optional CUDA event the DECODE of this batch waits for (its H2D copies do not).  The decoder's
        kernels are persistent launches of one CTA per SM with a static tile assignment: a kernel of another library that
        holds even a few SMs while they run (an NCCL collective spinning for its peer) makes every launch take two rounds.
        A caller that starts such a kernel per batch passes the event recorded behind it, so that collectives and decodes
        alternate on the SMs while the copies still overlap both (bench.py at N > 1: 18.4 -> 8.5 ms per step)."""
        for t in (z_host, g_host, out_host):   # out_host may be None when on_device consumes the result
            if t is not None and not (t.device.type == "cpu" and t.is_pinned()):
                raise RuntimeError("HostPipeline.submit: host tensors must be pinned CPU tensors")
        slot = self._n % len(self.streams)
        s = self.streams[slot]
        if self.done[slot] is not None:
            self.done[slot].synchronize()      # the slot's previous batch (and its host buffers) are finished
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.no_grad(), torch.cuda.stream(s):
            z = z_host.to(self.device, non_blocking=True)
            g = None if g_host is None else g_host.to(self.device, non_blocking=True)
            if decode_after is not None:
                s.wait_event(decode_after)
            y = self.G(z, g)
            if on_device is not None:
                on_device(y, s)
            if out_host is not None:
                out_host.copy_(y, non_blocking=True)
            # the caching allocator may hand these blocks to another stream as soon as they are dropped
            for t in (z, g, y):
                if t is not None:
                    t.record_stream(s)
            ev = torch.cuda.Event()
            ev.record(s)
        self.done[slot] = ev
        self._n += 1
        return slot

    def wait(self, ticket):
        if self.done[ticket] is not None:
            self.done[ticket].synchronize()

    def wait_all(self):
        for ev in self.done:
            if ev is not None:
                ev.synchronize()

    def join(self, stream=None):
        """Make `stream` (default: the current one) wait for everything submitted so far, without blocking the host."""
        stream = stream or torch.cuda.current_stream(self.device)
        for ev in self.done:
            if ev is not None:
                stream.wait_event(ev)
