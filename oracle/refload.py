"""Import the UNMODIFIED reference modules (networks, net_utils, kbnet_model) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Looks for the reference sources in baseline/_ref/src (installed by tools/install_reference.py; travels to the GPU box)
and then in /root/reference/src (build container).  The reference imports matplotlib in log_utils.py:20 for its
TensorBoard colour maps; matplotlib is not installed in this image, so empty stub modules are put into sys.modules
(nothing on the forward path touches them).  Returns None when no copy of the reference is available.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, 'baseline', '_ref', 'src'), '/root/reference/src')


def reference_dir():
    for d in CANDIDATES:
        if os.path.isfile(os.path.join(d, 'kbnet_model.py')):
            return d
    return None


def load():
    """-> (networks, net_utils, kbnet_model, directory) or None"""
    d = reference_dir()
    if d is None:
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    for m in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import networks, net_utils, kbnet_model  # noqa: E401
    return networks, net_utils, kbnet_model, d


def build_reference_model(cfg, device, f_img, f_dep, f_dec, seed=0):
    """kbnet_model.KBNetModel exactly as src/kbnet.py:785-800 builds it (keyword arguments), seeded."""
    import torch
    mods = load()
    if mods is None:
        return None
    kbnet_model = mods[2]
    torch.manual_seed(seed)
    model = kbnet_model.KBNetModel(
        input_channels_image=3, input_channels_depth=2,
        min_pool_sizes_sparse_to_dense_pool=cfg['min_pools'], max_pool_sizes_sparse_to_dense_pool=cfg['max_pools'],
        n_convolution_sparse_to_dense_pool=3, n_filter_sparse_to_dense_pool=8,
        n_filters_encoder_image=list(f_img), n_filters_encoder_depth=list(f_dep), resolutions_backprojection=[0, 1, 2, 3],
        n_filters_decoder=list(f_dec), deconv_type='up', weight_initializer='xavier_normal', activation_func='leaky_relu',
        min_predict_depth=cfg['min_pred'], max_predict_depth=cfg['max_pred'], device=device)
    if torch.device(device).type == 'cpu':
        # src/kbnet_model.py:408-415 wraps every network in torch.nn.DataParallel; on a box WITH GPUs the wrapper insists on
        # cuda:0 parameters.  An empty device_ids list is DataParallel's own CPU pass-through (what it does on a GPU-less
        # box) -- a setting on the torch wrapper object, the reference code is untouched.
        for net in (model.sparse_to_dense_pool, model.encoder, model.decoder):
            net.device_ids = []
    return model
