from .geom2d import ray_angles, dist_to_coord, polygons_to_label, polygons_to_label_coord
