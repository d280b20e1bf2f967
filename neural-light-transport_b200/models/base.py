"""Mirror of nlt/models/base.py:26-140 (trackability only; no TF)."""
import losses
from networks import base as basenet


class Model:
    def __init__(self, config):
        self.config = config
        self.net = {'main': basenet.Network()}  # NOTE: insert trainable
        # networks of your model into this dictionary
        self.trainable_registered = False
        self.wloss = self._init_loss()

    def _init_loss(self):
        wloss = []
        loss_str = self.config.get('DEFAULT', 'loss')
        for x in loss_str.split(','):
            loss_name, weight = self._parse_loss_and_weight(x)
            if loss_name == 'lpips':
                loss = losses.LPIPS(per_ch=False)
            elif loss_name == 'l1':
                loss = losses.L1()
            elif loss_name == 'l2':
                loss = losses.L2()
            elif loss_name == 'ssim':
                loss = losses.SSIM(1 - 0)
            else:
                raise NotImplementedError(loss_name)
            wloss.append((weight, loss))
        return wloss

    @staticmethod
    def _parse_loss_and_weight(weight_loss_str):
        """Handles strings like '1e+2lpips' or 'l1,10barron': the longest
        prefix that parses as a float is the weight (base.py:63-77)."""
        for i in range(len(weight_loss_str), -1, -1):
            try:
                weight = float(weight_loss_str[:i])
            except ValueError:
                continue
            return weight_loss_str[i:], weight
        return weight_loss_str, 1.

    def register_trainable(self):
        """Adds aliases `net_<name>_layer<i>` directly under `self` for every
        layer of every net (base.py:79-101); these names are the reference's
        checkpoint keys."""
        registered = []
        pref = 'net_'
        for net_name, net in self.net.items():
            attr_name = pref + net_name
            assert attr_name.isidentifier(), (
                "Prepending '{pref}' to your network name '{net}' doesn't "
                "make a valid identifier; change your network name").format(
                    pref=pref, net=net_name)
            for layer_i, layer in enumerate(net.layers):
                attr_name_full = attr_name + '_layer%d' % layer_i
                assert not hasattr(self, attr_name_full), \
                    "Can't register `%s` because it is already an attribute" \
                    % attr_name_full
                setattr(self, attr_name_full, layer)
                registered.append(attr_name_full)
        self._registered = registered
        self.trainable_registered = True

    @staticmethod
    def _validate_mode(mode):
        allowed_modes = ('train', 'vali', 'test')
        if mode not in allowed_modes:
            raise ValueError(mode)

    def __call__(self, batch, mode=None, **kwargs):
        return self.call(batch, mode, **kwargs)

    def call(self, batch, mode):
        raise NotImplementedError

    def compute_loss(self, pred, gt, **kwargs):
        raise NotImplementedError

    def vis_batch(self, data_dict, outdir, mode, dump_raw_to=None):
        raise NotImplementedError

    def compile_batch_vis(self, batch_vis_dirs, outpref, mode):
        raise NotImplementedError
