"""FrameBuffer (engine/src/framebuffer.rs:6-82).

`buffer` is an (height, width, 3) array: float32 for the production FP32 kernels, float64 in
RM_FP64 validation mode (the reference stores f64 Vec3f rows, framebuffer.rs:9)."""
import numpy as np


class FrameBuffer:
    def __init__(self, width, height, dtype=np.float32):
        self.width = int(width)
        self.height = int(height)
        self.buffer = np.zeros((self.height, self.width, 3), dtype=dtype)

    def to_vec(self):
        """framebuffer.rs:40-55,80-82: (255 * clamp(f, 0, 1)) as u8, truncating."""
        b = self.buffer
        one = b.dtype.type(1)
        return (b.dtype.type(255) * np.minimum(np.maximum(b, 0), one)).astype(np.uint8)

    def normalize(self):
        """framebuffer.rs:58-77: divide by the global channel maximum via scale(1/max)."""
        max_val = self.buffer.max() if self.buffer.size else 0
        max_val = max(max_val, 0)
        if max_val > 0:
            self.buffer *= self.buffer.dtype.type(1) / self.buffer.dtype.type(max_val)

    def write_ppm(self, filename):
        """framebuffer.rs:26-38."""
        with open(filename, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (self.width, self.height))
            f.write(self.to_vec().tobytes())
        return 0


def create_frame_buffer(width, height, dtype=np.float32):
    return FrameBuffer(width, height, dtype)
