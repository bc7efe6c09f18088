"""NOT a replacement of the reference's `fnet` package: only the plugin module `fnet.nn_modules.RepMode` (the name the
reference resolves at fnet/fnet_model.py:52) and a behaviour-level mirror of the `fnet.fnet_model.Model` host wrapper live
here, so that this repository's tests and bench can drive the path the reference's way on a box without the reference.
To use the path FROM the reference tree, replace the reference's one file fnet/nn_modules/RepMode.py;
the B200 implementation is in repmode_b200/."""
