"""B200-native replacement of the reference's `torch_utils` package (same module paths:
torch_utils.ops.{upfirdn2d,bias_act,conv2d_resample,conv2d_gradfix,fma}, torch_utils.custom_ops,
torch_utils.misc), see SURVEY.md section 8(b)."""
