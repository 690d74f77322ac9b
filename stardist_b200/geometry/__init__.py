from .geom2d import ray_angles, dist_to_coord, polygons_to_label, polygons_to_label_coord
from .geom3d import polyhedron_to_label, dist_to_coord3D
