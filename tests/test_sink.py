"""Frame hand-off to the encoder (animateportrait_b200/sink.py): ordering, ring reuse, error
surfacing on CPU; device frames through the pinned ring on the GPU."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from animateportrait_b200.sink import FrameSink


def _frames(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, 256, 256, 3), dtype=torch.uint8, generator=g)


def test_batches_arrive_in_order_through_a_small_ring(tmp_path):
    path = tmp_path / "frames.rgb"
    batches = [_frames(n, i) for i, n in enumerate((3, 1, 7, 2, 5))]
    with open(path, "wb") as f:
        sink = FrameSink(f, slots=2, max_batch=4)   # 7- and 5-frame batches are split over ring slots
        for b in batches:
            sink.put(b)
        assert sink.close() == 18
    got = np.fromfile(path, dtype=np.uint8).reshape(18, 256, 256, 3)
    assert np.array_equal(got, torch.cat(batches).numpy())


def test_frames_reach_an_encoder_process_over_a_pipe(tmp_path):
    """Any process that reads raw rgb24 frames on stdin stands in for ffmpeg (not installed here)."""
    out = tmp_path / "piped.rgb"
    proc = subprocess.Popen([sys.executable, "-c", "import sys,shutil; shutil.copyfileobj(sys.stdin.buffer, open(sys.argv[1],'wb'))",
                             str(out)], stdin=subprocess.PIPE)
    sink = FrameSink(proc.stdin, process=proc)
    fr = _frames(6, 9)
    sink.put(fr)
    assert sink.close() == 6
    assert np.array_equal(np.fromfile(out, dtype=np.uint8).reshape(6, 256, 256, 3), fr.numpy())


def test_the_reference_encode_command_is_kept_with_raw_video_input():
    cmd = FrameSink.ffmpeg_command("out.mp4", 62.5)
    assert cmd[:2] == ["ffmpeg", "-loglevel"] and "rawvideo" in cmd and "rgb24" in cmd and "256x256" in cmd
    # main_end2end_module2.py:123: -c:v libx264 -y -vf format=yuv420p <video_name>
    assert cmd[-6:] == ["-c:v", "libx264", "-y", "-vf", "format=yuv420p", "out.mp4"]


def test_bad_frames_and_writer_errors_surface():
    sink = FrameSink(io.BytesIO())
    with pytest.raises(RuntimeError, match="uint8"):
        sink.put(torch.zeros(1, 256, 256, 3))
    sink.close()

    class Broken(io.RawIOBase):
        def write(self, b):
            raise OSError("disk full")

    sink = FrameSink(Broken())
    sink.put(_frames(1, 0))
    with pytest.raises(RuntimeError, match="frame sink failed"):
        sink.close()


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_clip_frames_go_from_the_device_to_the_stream(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import animateportrait_b200 as ap
    from animateportrait_b200 import synth
    from animateportrait_b200.clip import ClipRenderer
    dev = torch.device("cuda", 0)
    T = 7
    net = ap.define_G(3, 1, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [0], div=3, disp=3)
    net.module.load_state_dict(synth.make_state_dict(1, seed=3, bias_std=0.3))
    photo, matte, static, src, seq, flow, ifmask = synth.make_clip(T, output_nc=1, seed=32)
    r = ClipRenderer(net, batch=3)
    r.set_photo(photo.to(dev), src.to(dev), matte.to(dev), static.to(dev))
    want = r.render(seq.to(dev), flow.to(dev), ifmask.to(dev)).cpu().numpy()
    path = tmp_path / "clip.rgb"
    with open(path, "wb") as f, FrameSink(f, slots=2, max_batch=3) as sink:
        assert r.render_to(sink, seq.to(dev), flow.to(dev), ifmask.to(dev)) == T
    got = np.fromfile(path, dtype=np.uint8).reshape(T, 256, 256, 3)
    assert np.array_equal(got, want)
