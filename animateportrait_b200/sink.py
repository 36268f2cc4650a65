"""Frame hand-off to the video encoder.

The reference writes nine PNG files per frame through PIL (Module2/util/visualizer.py:16-52, called from
Module2/test.py:63-65), copies the `fake_B` ones into a second directory and only then starts
`ffmpeg -framerate <fps> -i %05d.png -c:v libx264 -vf format=yuv420p out.mp4` on the image sequence
(main_end2end_module2.py:111-126).  Here the uint8 HWC frames that `ap_netg_compose` leaves on the device
(include/ap_netg.h) go through a ring of pinned host buffers straight into the encoder's stdin as raw video:

    device frames --(D2H on a copy stream, one ring slot per batch)--> pinned ring --(writer thread)--> pipe

B200 has no NVENC, so the encoder itself stays a host process: `FrameSink.ffmpeg(...)` builds the reference's libx264
command with `-f rawvideo -pix_fmt rgb24 -s 256x256 -framerate <fps> -i -` in place of the PNG sequence; any writable
binary file object works as well (tests use a plain file).  The render loop never waits for the disk or the encoder
unless the whole ring is in flight.
"""
from __future__ import annotations

import queue
import shutil
import subprocess
import threading
from typing import IO, List, Optional

import torch


class FrameSink:
    """Ordered, asynchronous writer of uint8 frame batches [n, H, W, 3] to a binary stream."""

    def __init__(self, stream: IO[bytes], height: int = 256, width: int = 256, slots: int = 4, max_batch: int = 64,
                 process: Optional[subprocess.Popen] = None):
        self.stream, self.process = stream, process
        self.h, self.w = int(height), int(width)
        self.max_batch = int(max_batch)
        pin = torch.cuda.is_available()
        self._ring: List[torch.Tensor] = [torch.empty((self.max_batch, self.h, self.w, 3), dtype=torch.uint8, pin_memory=pin)
                                          for _ in range(max(2, int(slots)))]
        self._free: "queue.Queue[int]" = queue.Queue()
        for i in range(len(self._ring)):
            self._free.put(i)
        self._todo: "queue.Queue" = queue.Queue()
        self._copy_stream = None
        self._error: Optional[BaseException] = None
        self.frames_written = 0
        self._thread = threading.Thread(target=self._writer, name="apnetg-frame-sink", daemon=True)
        self._thread.start()

    # ---- construction helpers -------------------------------------------------------------------------
    @staticmethod
    def ffmpeg_command(path: str, fps: float, height: int = 256, width: int = 256, exe: str = "ffmpeg") -> List[str]:
        """The reference's encode command (main_end2end_module2.py:123) fed with raw frames instead of PNG files."""
        return [exe, "-loglevel", "panic", "-f", "rawvideo", "-pix_fmt", "rgb24", "-s", f"{width}x{height}", "-framerate",
                str(fps), "-i", "-", "-c:v", "libx264", "-y", "-vf", "format=yuv420p", path]

    @classmethod
    def ffmpeg(cls, path: str, fps: float = 62.5, height: int = 256, width: int = 256, **kw) -> "FrameSink":
        exe = shutil.which("ffmpeg")
        if exe is None:
            raise RuntimeError("ffmpeg is not installed on this host; pass FrameSink a file object (raw rgb24 frames) instead")
        proc = subprocess.Popen(cls.ffmpeg_command(path, fps, height, width, exe), stdin=subprocess.PIPE)
        return cls(proc.stdin, height, width, process=proc, **kw)

    # ---- producer side --------------------------------------------------------------------------------
    def put(self, frames: torch.Tensor) -> None:
        """Queue a batch [n, H, W, 3] uint8 (device or host); returns as soon as the copy is enqueued.  Batches are
        written in the order they were put."""
        if self._error is not None:
            raise RuntimeError("frame sink writer failed") from self._error
        if frames.dtype != torch.uint8 or frames.dim() != 4 or tuple(frames.shape[1:]) != (self.h, self.w, 3):
            raise RuntimeError(f"frames: expected uint8 [n,{self.h},{self.w},3], got {frames.dtype} {tuple(frames.shape)}")
        for s in range(0, frames.shape[0], self.max_batch):
            part = frames[s:s + self.max_batch]
            slot = self._free.get()                     # blocks only when every ring slot is in flight
            dst = self._ring[slot][:part.shape[0]]
            ev = None
            if part.is_cuda:
                dev = part.device
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(dev)
                self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(self._copy_stream):
                    dst.copy_(part, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                part.record_stream(self._copy_stream)
            else:
                dst.copy_(part)
            self._todo.put((slot, part.shape[0], ev))

    def _writer(self) -> None:
        while True:
            job = self._todo.get()
            if job is None:
                return
            slot, n, ev = job
            try:
                if self._error is None:
                    if ev is not None:
                        ev.synchronize()
                    self.stream.write(memoryview(self._ring[slot][:n].numpy()).cast("B"))
                    self.frames_written += n
            except BaseException as e:  # surfaced by the next put() / close()
                self._error = e
            finally:
                self._free.put(slot)

    def close(self) -> int:
        """Flush, close the stream (and wait for the encoder process); returns the number of frames written."""
        self._todo.put(None)
        self._thread.join()
        try:
            self.stream.flush()
            if self.process is not None:
                self.stream.close()
                rc = self.process.wait()
                if rc != 0 and self._error is None:
                    self._error = RuntimeError(f"encoder exited with status {rc}")
        except BaseException as e:
            self._error = self._error or e
        if self._error is not None:
            raise RuntimeError("frame sink failed") from self._error
        return self.frames_written

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
