# Single-class (Car) SA-SSD inference config in the reference's config dialect
# (same keys/values as the reference's configs/car_cfg.py for model, test_cfg and the
# data-side voxel/anchor generators; training-only sections are omitted).  The
# reference's own file also loads unchanged through sassd_b200.Config.fromfile.
model = dict(
    type='SingleStageDetector',
    backbone=dict(type='SimpleVoxel', num_input_features=4, use_norm=True, num_filters=[32, 64],
                  with_distance=False),
    neck=dict(type='SpMiddleFHD', output_shape=[40, 1600, 1408], num_input_features=4,
              num_hidden_features=64 * 5),
    bbox_head=dict(type='SSDRotateHead', num_class=1, num_output_filters=256, num_anchor_per_loc=2,
                   use_sigmoid_cls=True, encode_rad_error_by_sin=True, use_direction_classifier=True,
                   box_code_size=7),
    extra_head=dict(type='PSWarpHead', grid_offsets=(0., 40.), featmap_stride=.4, in_channels=256,
                    num_class=1, num_parts=28),
)
train_cfg = None
test_cfg = dict(
    rpn=dict(nms_across_levels=False, nms_pre=2000, nms_post=100, nms_thr=0.7, min_bbox_size=0),
    extra=dict(score_thr=0.3, nms=dict(type='nms', iou_thr=0.1), max_per_img=100),
)
_generator = dict(type='VoxelGenerator', voxel_size=[0.05, 0.05, 0.1],
                  point_cloud_range=[0., -40., -3., 70.4, 40., 1.], max_num_points=5, max_voxels=20000)
_anchor = dict(type='AnchorGeneratorStride', anchor_strides=[0.4, 0.4, 1.0],
               anchor_offsets=[0.2, -39.8, -1.78], rotations=[0, 1.57])
data = dict(
    val=dict(class_names=['Car'], generator=_generator,
             anchor_generator=dict(Car=dict(_anchor, sizes=[1.6, 3.9, 1.56])),
             anchor_area_threshold=1, out_size_factor=8, test_mode=True),
)
