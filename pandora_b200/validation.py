"""validation step: mirror of AbstractValidation / CrossCheckingAccurate (src/pandora/validation/validation.py:43-371)
and of the right-disparity computation of the fast mode (state_machine.py:436-448), on the device."""
from __future__ import annotations

from typing import Dict

import numpy as np

from ._common import ConfigError, device_var, device_volume, get_engine, store_var
from .dataset import Dataset


def right_disparity_fast(cv, invalid_disparity: float = -9999.0):
    """state_machine.py:438-448 for ``cross_checking_fast``: the right disparity map is the WTA of the reversed left
    volume.  The right volume is never materialised (``pb200_wta_right``).  Returns the right disparity dataset."""
    eng = get_engine()
    cv_t = device_volume(eng, cv)
    disps = np.asarray(cv.coords["disp"].data)
    dmax = int(round(float(disps[-1])))
    is_max = cv.attrs.get("type_measure") == "max"
    disp_t, flags = eng.wta_right(cv_t, -dmax, is_max, float(invalid_disparity))
    H, W = (int(s) for s in disp_t.shape)
    out = Dataset(coords={"row": cv.coords["row"].data, "col": cv.coords["col"].data}, attrs=dict(cv.attrs))
    store_var(out, "disparity_map", disp_t)
    out["disparity_interval"] = (("disparity",), -np.asarray(disps)[[-1, 0]])
    import torch  # noqa: PLC0415

    mask = torch.where(flags != 0, 0x3C3, 0).to(torch.int16)      # disparity.py:470-474 on an empty right validity mask
    store_var(out, "validity_mask", mask, dtype="uint16")
    return out


class AbstractValidation:
    validation_methods_avail: Dict[str, type] = {}

    def __new__(cls, **cfg):
        if cls is AbstractValidation:
            method = cfg["validation_method"]                        # KeyError when absent, like the reference
            try:
                return super().__new__(cls.validation_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No validation method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str, *aliases):
        def decorator(subclass):
            for name in (short_name, *aliases):
                cls.validation_methods_avail[name] = subclass
            return subclass

        return decorator


@AbstractValidation.register_subclass("cross_checking_accurate", "cross_checking_fast")
class CrossCheckingAccurate(AbstractValidation):
    _THRESHOLD = 1.0

    def __init__(self, **cfg):
        self.cfg = self.check_conf(**cfg)
        self._threshold = self.cfg["cross_checking_threshold"]
        self._method = self.cfg["validation_method"]

    def check_conf(self, **cfg) -> dict:
        cfg.setdefault("cross_checking_threshold", self._THRESHOLD)
        for key in cfg:
            if key not in ("validation_method", "cross_checking_threshold", "interpolated_disparity"):
                raise ConfigError(f"Unknown key {key!r} in the validation configuration")
        if not isinstance(cfg["cross_checking_threshold"], (int, float)):
            raise ConfigError("cross_checking_threshold must be an int or a float")
        return cfg

    def desc(self) -> None:
        print("Cross-checking method")

    def disparity_checking(self, dataset_left, dataset_right, img_left=None, img_right=None, cv=None):
        """validation.py:226-371: occlusion / mismatch bits in the left validity mask and the left-right distance as
        the ``confidence_from_left_right_consistency`` indicator."""
        eng = get_engine()
        dl, dr = device_var(eng, dataset_left, "disparity_map"), device_var(eng, dataset_right, "disparity_map")
        mask_t = device_var(eng, dataset_left, "validity_mask", "uint16").clone()
        if "disparity_interval" in dataset_left:
            dmin, dmax = (int(v) for v in np.asarray(dataset_left["disparity_interval"].data))
        else:
            dmin, dmax = (int(v) for v in dataset_left.attrs["disparity_interval"])
        offset = int(dataset_left.attrs.get("offset_row_col", 0))
        conf = eng.cross_checking(dl, mask_t, dr, float(self._threshold), dmin, dmax, offset)
        store_var(dataset_left, "validity_mask", mask_t, dtype="uint16")
        dataset_left.attrs["validation"] = self._method
        from .cost_volume_confidence import AbstractCostVolumeConfidence  # noqa: PLC0415

        dataset_left, _ = AbstractCostVolumeConfidence.allocate_confidence_map("left_right_consistency", conf.cpu().numpy(), dataset_left, cv)
        return dataset_left
