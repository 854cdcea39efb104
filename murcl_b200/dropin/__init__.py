"""Drop-in replacements for the reference's hot-path modules, one file per upstream file:

    upstream                      here
    models/abmil.py          ->   murcl_b200.dropin.abmil
    models/clam.py           ->   murcl_b200.dropin.clam
    models/dsmil.py          ->   murcl_b200.dropin.dsmil
    models/rlmil.py          ->   murcl_b200.dropin.rlmil
    models/cl.py             ->   murcl_b200.dropin.cl
    utils/losses.py          ->   murcl_b200.dropin.losses
    utils/datasets.py        ->   murcl_b200.dropin.datasets   (get_feats, mixup only)

Same class / function names, constructor arguments, forward signatures, return arity, error
behaviour and state-dict keys; every computation runs in libmurcl_b200.so.  INTEGRATION.md shows the
import lines a maintainer changes in train_MuRCL.py / train_RLMIL.py.
"""
