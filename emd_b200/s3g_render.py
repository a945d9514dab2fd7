"""S3Gaussian + EMD render step -- host-side mirror of ``render`` (``S3Gaussian/gaussian_renderer/__init__.py:27-303``)
and of the image / deformation terms of the training loss (``S3Gaussian/train.py:207-366``), composed from this
package's kernels: HexPlane gather (K1e) -> EMD deformation MLP (K1d) -> activations -> 1..3 ``diff_gauss`` rasterizer
passes (K2', K3-K7; RGB + depth + alpha, then the coarse and the fine feature maps) -> sky blend -> fused image losses.

``render`` keeps the reference's signature and result keys, so ``train.py``'s consumers read it unchanged
(``render_pkg["render" | "viewspace_points" | "visibility_filter" | "radii" | "depth" | "weight" | "ddict" | "feat_c" |
"feat_f" | "sky_color"]``); the extra key ``"color"`` is the rasterizer's own colour before the sky blend, which the
fused loss kernel blends itself.  ``pc`` is duck-typed like ``GaussianModel``; :class:`S3GGaussians` is this package's
implementation of that interface (parameters keep their ``GaussianModel`` names).

Not mirrored (outside BASELINE.json configs[2], raise): ``compute_cov3D_python``, ``convert_SHs_python`` /
``override_color``, ``combine_dynamic_static``, ``return_decomposition`` (eval-time visualisation renders).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import torch
from torch import Tensor

from .diff_gauss_api import GaussianRasterizationSettings, GaussianRasterizer
from .emd_s3g import REG_KEYS, S3GDeformation
from .losses import ImageLossConfig, s3g_image_losses
from .sh_ops import activate_geometry


@dataclass
class S3GOptions:
    """The ``BaseOptions`` attributes ``render`` / the loss read (``S3Gaussian/arguments/gaussian_options.py``), with the
    values of the reference's run scripts (``--no_ds --no_dr --no_fine_hexplane_features``)."""
    debug: bool = False
    compute_cov3D_python: bool = False
    convert_SHs_python: bool = False
    combine_dynamic_static: bool = False
    no_coarse_deform: bool = False
    no_fine_deform: bool = False
    feat_head: bool = True
    no_dx: bool = False
    no_do: bool = False
    no_dshs: bool = False
    lambda_dssim: float = 0.2
    lambda_depth: float = 0.5
    lambda_dx: float = 0.001
    lambda_do: float = 0.001
    lambda_dshs: float = 0.001
    lambda_f2c: float = 0.0
    lambda_feat: float = 0.001
    lambda_sky: float = 0.05
    load_sky_mask: bool = True


@dataclass
class S3GCamera:
    """The ``Camera`` attributes ``render`` reads (``S3Gaussian/scene/cameras.py:20-75``)."""
    FoVx: float
    FoVy: float
    image_height: int
    image_width: int
    world_view_transform: Tensor   # [4,4], transposed (row-vector convention)
    full_proj_transform: Tensor    # [4,4], transposed
    camera_center: Tensor          # [3]
    time: float = 0.0
    cam_no: int = 0
    time_diff: float = 0.0


def make_camera(yaw_deg: float, width: int, height: int, time: float = 0.0, cam_no: int = 0, znear: float = 0.01,
                zfar: float = 100.0, device="cpu") -> S3GCamera:
    """A ``Camera`` on the synthetic street (``cameras.py:55-66``, ``graphics_utils.py:72-92``: getWorld2View2 +
    getProjectionMatrix, both transposed)."""
    from . import scenes
    c2w, K = scenes.camera(yaw_deg, width, height)
    fovx = 2 * math.atan(width / (2 * float(K[0, 0])))
    fovy = 2 * math.atan(height / (2 * float(K[1, 1])))
    world_view = torch.linalg.inv(c2w).transpose(0, 1).contiguous()
    tx, ty = math.tan(fovx / 2), math.tan(fovy / 2)
    top, right = ty * znear, tx * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    full = world_view @ P.transpose(0, 1)
    center = torch.linalg.inv(world_view)[3, :3]
    return S3GCamera(fovx, fovy, height, width, world_view.to(device), full.contiguous().to(device),
                     center.contiguous().to(device), time, cam_no)


class S3GGaussians:
    """``GaussianModel`` (``S3Gaussian/scene/gaussian_model.py``) as far as ``render`` uses it: the raw parameters under
    their reference names, the activations (``:40-48``), ``_deformation`` (``deform_network``), ``_sky_model``."""

    def __init__(self, params: Dict[str, Tensor], deformation: S3GDeformation, sky_model: Optional[Callable] = None,
                 active_sh_degree: int = 3):
        self._xyz, self._scaling, self._rotation = params["_xyz"], params["_scaling"], params["_rotation"]
        self._opacity, self._features_dc, self._features_rest = params["_opacity"], params["_features_dc"], params["_features_rest"]
        self._embedding = params["_embedding"]
        self.deform = deformation
        self.sky = sky_model
        self.active_sh_degree = active_sh_degree
        self._deformation_table = None
        self.fused_residuals = self._features_rest.shape[1] == 15     # the fused kernel is built for SH degree 3
        self.scaling_activation = torch.exp
        self.rotation_activation = torch.nn.functional.normalize
        self.opacity_activation = torch.sigmoid

    get_xyz = property(lambda self: self._xyz)
    get_embedding = property(lambda self: self._embedding)

    @property
    def get_features(self) -> Tensor:
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    def parameters(self):
        ps = [self._xyz, self._scaling, self._rotation, self._opacity, self._features_dc, self._features_rest, self._embedding]
        ps += [v for v in self.deform.w.values() if v.requires_grad]
        if self.deform.grid is not None:
            ps += [p for p in self.deform.grid.parameters() if p.requires_grad]
        return ps

    def _deformation(self, point, scales, rotations, opacity, shs, time, embeddings, iteration, cam_no, time_diff=None,
                     is_train=False):
        """``deform_network.forward`` (``deformation.py:484-527``); ``time`` arrives as the [N,1] repeat of one value."""
        t = time.reshape(-1)[0] if isinstance(time, Tensor) else time
        if shs is None:     # fused residual application: concatenates features_dc / features_rest itself
            return self.deform(point, scales, rotations, opacity, None, t, embeddings, iteration, cam_no,
                               shs_parts=(self._features_dc, self._features_rest))
        return self.deform(point, scales, rotations, opacity, shs, t, embeddings, iteration, cam_no)

    def activate(self, scales, rotations, opacity):
        """exp / normalize / sigmoid (``gaussian_renderer/__init__.py:99-101``) in one fused kernel pass."""
        op, sc, qn = activate_geometry(opacity, scales, rotations)
        return sc, qn, op[:, None]

    def _sky_model(self, viewpoint_camera, acc=None, is_train=False) -> Tensor:
        if self.sky is None:
            return torch.zeros(3, viewpoint_camera.image_height, viewpoint_camera.image_width, device=self._xyz.device)
        return self.sky(viewpoint_camera, acc, is_train)


def render(args: S3GOptions, viewpoint_camera, pc, bg_color: Tensor, scaling_modifier: float = 1.0, override_color=None,
           stage: str = "fine", return_decomposition: bool = False, return_dx: bool = False, render_feat: bool = False,
           iter: Optional[int] = None, is_train: bool = False) -> Dict[str, Tensor]:  # noqa: A002 (reference name)
    if args.compute_cov3D_python or args.convert_SHs_python or override_color is not None:
        raise NotImplementedError("emd_b200.s3g_render: compute_cov3D_python / convert_SHs_python / override_color are not "
                                  "mirrored (off in the reference's configuration, gaussian_options.py:63-64)")
    if args.combine_dynamic_static or return_decomposition:
        raise NotImplementedError("emd_b200.s3g_render: combine_dynamic_static / return_decomposition are eval-time "
                                  "visualisation paths outside the training step")
    means3D = pc.get_xyz
    dev = means3D.device
    # the screen-space gradient holder of the reference (__init__.py:36): its .grad[:, :2] is the densification statistic
    screenspace_points = torch.zeros_like(means3D, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except RuntimeError:
        pass
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform.to(dev),
        projmatrix=viewpoint_camera.full_proj_transform.to(dev), sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center.to(dev), prefiltered=False, debug=args.debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    # S3GGaussians applies the residuals in a fused kernel that reads features_dc / features_rest directly (shs=None)
    fused = "fine" in stage and getattr(pc, "fused_residuals", False)
    opacity, scales, rotations = pc._opacity, pc._scaling, pc._rotation
    shs = None if fused else pc.get_features
    ddict = None
    if "coarse" in stage:
        means3D_final, scales_final, rotations_final, opacity_final, shs_final = means3D, scales, rotations, opacity, shs
    elif "fine" in stage:
        time = torch.tensor(viewpoint_camera.time, device=dev).repeat(1, 1)   # the kernels take the scalar, not [N,1]
        means3D_final, scales_final, rotations_final, opacity_final, shs_final, ddict = pc._deformation(
            means3D, scales, rotations, opacity, shs, time, pc.get_embedding, iter, viewpoint_camera.cam_no,
            viewpoint_camera.time_diff, is_train=is_train)
    else:
        raise NotImplementedError
    if hasattr(pc, "activate"):
        scales_final, rotations_final, opacity_final = pc.activate(scales_final, rotations_final, opacity_final)
    else:
        scales_final = pc.scaling_activation(scales_final)
        rotations_final = pc.rotation_activation(rotations_final)
        opacity_final = pc.opacity_activation(opacity_final)

    def raster(shs_=None, colors_=None):
        return rasterizer(means3D=means3D_final, means2D=screenspace_points, shs=shs_, colors_precomp=colors_,
                          opacities=opacity_final, scales=scales_final, rotations=rotations_final, cov3Ds_precomp=None,
                          extra_attrs=None)

    rendered_image, depth, normal, weight, radii, _ = raster(shs_=shs_final)
    result = {"render": rendered_image, "color": rendered_image, "viewspace_points": screenspace_points,
              "visibility_filter": radii > 0, "radii": radii, "depth": depth, "weight": weight, "normal": normal}
    if render_feat and "fine" in stage:
        result["feat_c"] = None if args.no_coarse_deform else raster(colors_=ddict["coarse"]["feat"])[0]
        result["feat_f"] = None if args.no_fine_deform else raster(colors_=ddict["fine"]["feat"])[0]
    if return_dx and "fine" in stage:
        result["ddict"] = ddict
    sky_color = pc._sky_model(viewpoint_camera, acc=weight, is_train=is_train)
    result["render"] = result["render"] * result["weight"] + sky_color * (1 - result["weight"])
    result["sky_color"] = sky_color
    return result


def training_losses(args: S3GOptions, render_pkg: Dict[str, Tensor], gt_image: Tensor, gt_depth: Tensor,
                    sky_mask: Optional[Tensor], gt_feat: Optional[Tensor] = None, stage: str = "fine") -> Dict[str, Tensor]:
    """The terms of ``S3Gaussian/train.py:226-363`` that depend on the render: L1, the deformation L1 regularisers
    (dx, do, dshs, coarse + fine, f2c), the feature-map L2 (``feat_head``), depth L2, D-SSIM, sky.  The four image terms
    come from ONE fused kernel pass over the rasterizer's colour / depth / alpha and the sky colour (it forms the sky
    blend itself); ``loss = sum(values)`` as in the reference.  Not here: the KNN embedding regulariser (CPU KNN,
    ``train.py:326-338``) and the plane TV terms (``time_smoothness_weight`` is 0 in the reference's options)."""
    cfg = ImageLossConfig.s3g(lambda_dssim=args.lambda_dssim, lambda_depth=args.lambda_depth,
                              lambda_sky=args.lambda_sky if (sky_mask is not None and args.load_sky_mask) else 0.0)
    img = s3g_image_losses(render_pkg["color"], render_pkg["depth"], render_pkg["weight"], render_pkg["sky_color"], gt_image,
                           gt_depth, sky_mask if args.load_sky_mask else None, cfg)
    out = dict(img)
    if "fine" in stage:
        dd = render_pkg["ddict"]
        for key, lam, off in (("dx", args.lambda_dx, args.no_dx), ("do", args.lambda_do, args.no_do),
                              ("dshs", args.lambda_dshs, args.no_dshs)):
            if off or lam == 0:
                continue
            terms = []
            for branch, off_b in (("coarse", args.no_coarse_deform), ("fine", args.no_fine_deform)):
                if off_b:
                    continue
                if "reg_sums" in dd:       # sum |.| from the fused residual kernel (emd_s3g.apply_residuals)
                    terms.append(dd["reg_sums"][REG_KEYS.index((branch, key))] * (lam / max(dd[branch][key].numel(), 1)))
                else:
                    terms.append(dd[branch][key].abs().mean() * lam)
            if key == "dx" and not args.no_fine_deform and not args.no_coarse_deform and args.lambda_f2c != 0:
                terms.append((dd["fine"]["dx"] - dd["coarse"]["dx"]).abs().mean() * args.lambda_f2c)
            out[key + "_loss"] = sum(terms)
        if args.feat_head and gt_feat is not None:
            terms = []
            if not args.no_coarse_deform:
                terms.append(((render_pkg["feat_c"] - gt_feat) ** 2).mean() * args.lambda_feat)
            if not args.no_fine_deform:
                terms.append(((render_pkg["feat_f"] - gt_feat) ** 2).mean() * args.lambda_feat)
            out["loss_feat"] = sum(terms)
    return out
