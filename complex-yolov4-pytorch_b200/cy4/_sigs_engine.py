"""ctypes signatures of the conv-engine entry points (include/cy4.h, "convolution stack")."""
import ctypes

c_f = ctypes.c_void_p
c_vp = ctypes.c_void_p
c_i = ctypes.c_int
c_i64 = ctypes.c_int64


class ConvDesc(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int32), ("Hi", ctypes.c_int32), ("Wi", ctypes.c_int32), ("Cin", ctypes.c_int32),
                ("Ho", ctypes.c_int32), ("Wo", ctypes.c_int32), ("Cout", ctypes.c_int32),
                ("ksize", ctypes.c_int32), ("stride", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("ldx", ctypes.c_int64), ("ldy", ctypes.c_int64),
                ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


PD = ctypes.POINTER(ConvDesc)


class PackItem(ctypes.Structure):
    _fields_ = [("w_oihw", ctypes.c_void_p), ("w_fprop", ctypes.c_void_p), ("w_dgrad", ctypes.c_void_p),
                ("Cout", ctypes.c_int32), ("Cin", ctypes.c_int32), ("ksize", ctypes.c_int32), ("cout_pad", ctypes.c_int32),
                ("cin_pad", ctypes.c_int32), ("tile_begin", ctypes.c_int32), ("fold_scale", ctypes.c_void_p)]


class UnpackItem(ctypes.Structure):
    _fields_ = [("dw_acc", ctypes.c_void_p), ("gw_oihw", ctypes.c_void_p), ("Cout", ctypes.c_int32), ("Cin", ctypes.c_int32),
                ("ksize", ctypes.c_int32), ("tile_begin", ctypes.c_int32)]

SIGS = {
    "cy4_set_option": (c_i, [ctypes.c_char_p, c_i]),
    "cy4_conv_fwd": (c_i, [PD, c_f, c_f, c_f, c_f, c_f, c_f, c_vp]),
    "cy4_conv_dgrad": (c_i, [PD, c_f, c_f, c_f, c_vp]),
    "cy4_conv_fwd_fused": (c_i, [PD, c_f, c_f, c_f, c_f, c_i, c_f, c_i64, c_vp]),
    "cy4_conv_dgrad_fused": (c_i, [PD, c_f, c_f, c_f, c_f, c_i64, c_f, c_f, c_i, c_f, c_f, c_vp]),
    "cy4_conv_wgrad": (c_i, [PD, c_f, c_f, c_f, c_vp]),
    "cy4_conv_wgrad_plan": (c_i, [PD, c_vp]),
    "cy4_pack_weight_fprop": (c_i, [c_f, c_i, c_i, c_i, c_i, c_f, c_vp]),
    "cy4_pack_weight_dgrad": (c_i, [c_f, c_i, c_i, c_i, c_f, c_vp]),
    "cy4_unpack_wgrad": (c_i, [c_f, c_i, c_i, c_i, c_i, ctypes.c_float, c_f, c_i, c_f, c_vp]),
    "cy4_pack_weights_batched": (c_i, [c_f, c_i, c_vp]),
    "cy4_unpack_wgrad_batched": (c_i, [c_f, c_i, c_f, c_vp]),
    "cy4_absmax_f32": (c_i, [c_f, c_i64, c_f, c_vp]),
    "cy4_make_scale": (c_i, [c_f, ctypes.c_float, c_f, c_vp]),
    "cy4_bn_finalize": (c_i, [c_f, c_f, ctypes.c_float, c_f, c_f, c_f, c_f, c_f, ctypes.c_float, ctypes.c_float, c_i, c_i,
                                c_f, c_f, c_f, c_f, c_vp]),
    "cy4_bn_train_act_fwd": (c_i, [c_f, c_i64, c_f, c_f, ctypes.c_float, c_f, c_f, c_f, c_f, c_f, ctypes.c_float, ctypes.c_float,
                                     c_f, c_f, c_f, c_f, c_i, c_f, c_i64, c_f, c_i64, c_i64, c_i, c_f, c_f, c_vp]),
    "cy4_conv_fwd_stats": (c_i, [PD, c_f, c_f, c_f, c_f, c_f, c_f, c_vp]),
    "cy4_bn_act_fwd": (c_i, [c_f, c_i64, c_f, c_f, c_i, c_f, c_i64, c_f, c_i64, c_i64, c_i, c_vp]),
    "cy4_bn_act_bwd_reduce": (c_i, [c_f, c_i64, c_f, c_i64, c_f, c_f, c_f, c_f, c_i, c_i64, c_i, c_f, c_f, c_vp]),
    "cy4_bn_bwd_fixup": (c_i, [c_f, c_f, c_f, c_f, c_i, c_vp]),
    "cy4_bn_act_bwd_apply": (c_i, [c_f, c_i64, c_f, c_i64, c_f, c_f, c_f, c_f, c_f, c_f, ctypes.c_float, c_i, c_i, c_i, c_f, c_i64,
                                     c_i64, c_i, c_vp]),
    "cy4_add_copy": (c_i, [c_f, c_i64, c_f, c_i64, c_f, c_i64, c_i64, c_i, c_vp]),
    "cy4_upsample2x_fwd": (c_i, [c_f, c_i64, c_f, c_i64, c_i, c_i, c_i, c_i, c_vp]),
    "cy4_upsample2x_bwd": (c_i, [c_f, c_i64, c_f, c_i64, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "cy4_maxpool_fwd": (c_i, [c_f, c_i64, c_f, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "cy4_maxpool_bwd": (c_i, [c_f, c_i64, c_f, c_i64, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "cy4_maxpool_fwd_idx": (c_i, [c_f, c_i64, c_f, c_i64, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "cy4_maxpool_bwd_idx": (c_i, [c_f, c_f, c_i64, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "cy4_f32_to_f16": (c_i, [c_f, c_i64, ctypes.c_float, c_f, c_f, c_i64, c_i64, c_i, c_i, c_vp]),
    "cy4_colsum_f32": (c_i, [c_f, c_i64, c_i64, c_i, ctypes.c_float, c_f, c_i, c_vp]),
    "cy4_stem_im2col": (c_i, [c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_vp]),
}

CONV_OUT_F32 = 1
CONV_STATS = 2
CONV_ACCUM = 4
CONV_A_MATRIX = 8
CONV_ZERO_ACC = 16
