"""Derives the committed golden fixture from the ONE output image the reference publishes: doc/vk_order_independent_transparency.png
(README.md:5), a 1920x1016 capture of the running sample with its GUI in the top-left corner.  The GUI panel states the
configuration: Interlock, 100 % transparent, alpha 0.2 + 0.3, tail blend on, 16 layers, MSAA 4x pixel shading, 1024 objects,
subdivision 16, scale 0.1 + 0.9, viewport 1920 x 1017 ("Aux image: 1920 x 1017"), A-buffer 499,875,840 bytes.

The fixture is that image with the GUI rectangle (x < 360, y < 400) blanked to black, losslessly re-encoded.
Run in the build container (needs /root/reference): python tests/golden/make_reference_screenshot.py"""
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/doc/vk_order_independent_transparency.png"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_screenshot_interlock16_msaa4.png")
GUI_W, GUI_H = 360, 400

if __name__ == "__main__":
    img = np.asarray(Image.open(SRC).convert("RGB")).copy()
    assert img.shape == (1016, 1920, 3), img.shape
    img[:GUI_H, :GUI_W] = 0
    Image.fromarray(img).save(DST, optimize=True)
    print("wrote", DST, os.path.getsize(DST), "bytes")
