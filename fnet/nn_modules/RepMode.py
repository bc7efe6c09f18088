"""Plugin module resolved by `importlib.import_module('fnet.nn_modules.' + nn_module).Net(opts)`
(reference fnet/fnet_model.py:52).  Re-exports the B200-native implementation under the reference's names."""
from repmode_b200.nn_modules import (MoDEConv, MoDEDecoderBlock, MoDEEncoderBlock, MoDESubNet2Conv,  # noqa: F401
                                     Net)
