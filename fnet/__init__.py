"""Drop-in `fnet` package surface for the reference drivers (main.py / eval.py): only the plugin module
`fnet.nn_modules.RepMode` and the `fnet.fnet_model.Model` host wrapper live here; the B200 implementation is
in repmode_b200/."""
