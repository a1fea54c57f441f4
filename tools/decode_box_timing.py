"""Times the stand-alone head transpose+decode (`decode_box`, K0/K0t) on the headline head levels.
    python tools/decode_box_timing.py
Bytes = read + write of the [B, A*(5+C), G, G] fp32 head."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import objectdetectionpl_b200 as od  # noqa: E402


def main():
    dev = "cuda:0"
    B, A, C = 64, 3, 80
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for G, mode in [(80, "yolo_exp"), (80, "yolov5"), (52, "none"), (40, "yolo_exp"), (20, "yolo_exp"), (13, "yolo_exp")]:
        head = torch.randn(B, A * (5 + C), G, G, device=dev)
        anc = torch.tensor([[10., 13.], [16., 30.], [33., 23.]], device=dev)
        od.decode_box(head, anc, 8.0, mode, num_anchors=A)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            od.decode_box(head, anc, 8.0, mode, num_anchors=A)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        nbytes = 2 * head.numel() * 4
        print(f"decode_box G={G} mode={mode}: {us:.1f} us, {nbytes / us / 1e3:.0f} GB/s ({nbytes / 1e6:.0f} MB read+write)")


if __name__ == "__main__":
    main()
