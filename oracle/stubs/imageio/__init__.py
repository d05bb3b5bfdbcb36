"""Stand-in for imageio (test_w.py:9,115 writes a GIF with it). TEST INFRASTRUCTURE ONLY."""


def mimsave(path, frames, *a, **k):
    with open(path, "wb") as f:
        f.write(b"GIF89a-stub:%d-frames" % len(frames))
