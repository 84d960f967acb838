"""Drop-in for the reference extension module `pointnet2_cuda`
(pointnet2_lib/pointnet2/src/pointnet2_api.cpp:11-23): same nine functions, same signatures,
backed by libws3d_ops.so.  Put this directory on sys.path (or call ws3d_b200.install_dropins())
and the reference's pointnet2_utils.py imports it unmodified."""
from ws3d_b200.native import (ball_query_wrapper, furthest_point_sampling_wrapper, gather_points_grad_wrapper,  # noqa: F401
                              gather_points_wrapper, group_points_grad_wrapper, group_points_wrapper,
                              three_interpolate_grad_wrapper, three_interpolate_wrapper, three_nn_wrapper)
