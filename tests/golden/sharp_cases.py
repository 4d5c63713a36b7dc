"""The inputs of the sharp pin (tools/sharp_fixtures.py writes them as PNG — no lossy codec between these pixels and
sharp — tools/sharp_golden.mjs runs the reference's own calls on them, tests/test_sharp_golden.py compares).

A case is (name, h, w, c, seed, kind, exif_orientation); pixels = tests/conftest.rand_image(h, w, c, seed, kind).
Classify cases stay small enough for the dump to carry whole buffers (grey, the three stencil responses, the blurred
image); preprocess cases are wide or tall STRIPS, so that they take real shrink factors (1.27 ... 4.4, the last through
vips_resize's box pre-shrink) while their resized pixels stay a few hundred KB."""

CLASSIFY_CASES = [
    ("c_1x1", 1, 1, 3, 1, "noise", 1),
    ("c_7x5", 7, 5, 3, 2, "noise", 1),
    ("c_33x257_smooth", 33, 257, 3, 3, "smooth", 1),
    ("c_128_noise", 128, 128, 3, 4, "noise", 1),
    ("c_97x513_edges", 97, 513, 3, 5, "edges", 1),
    ("c_120x200_grey", 120, 200, 1, 6, "smooth", 1),
    ("c_64_rgba", 64, 64, 4, 7, "noise", 1),
    ("c_200x240_smooth", 200, 240, 3, 8, "smooth", 1),
]

PREPROCESS_CASES = [
    ("p_2600x32_noise", 32, 2600, 3, 11, "noise", 1),          # shrink 1.27
    ("p_32x2600_smooth", 2600, 32, 3, 12, "smooth", 1),        # tall
    ("p_4096x32_noise", 32, 4096, 3, 13, "noise", 1),          # shrink 2.0 exactly
    ("p_6000x40_smooth", 40, 6000, 3, 14, "smooth", 1),        # shrink 2.93 (24 MP photos)
    ("p_9000x64_noise", 64, 9000, 3, 15, "noise", 1),          # shrink 4.39: box pre-shrink by 2, then 2.2
    ("p_2600x48_rot6", 48, 2600, 3, 16, "smooth", 6),          # EXIF 6: the pre-rotation-dims quirk
    ("p_300x200_keep", 200, 300, 3, 17, "noise", 1),           # no resize
    ("p_2500x40_grey", 40, 2500, 1, 18, "smooth", 1),
    ("p_2500x40_rgba", 40, 2500, 4, 19, "noise", 1),           # sharp premultiplies: the documented deviation
]

ALL_CASES = CLASSIFY_CASES + PREPROCESS_CASES
