import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util
pts, nl = util.scene("pcisph", "asshipped")
o = util.make_oracle("pcisph", pts, nl)
m = util.make_engine("pcisph", pts, nl)
for s in range(3):
    m.particle_data.pos.from_numpy(o.field("pos")); m.particle_data.vel.from_numpy(o.field("vel"))
    o.call("update_grid"); m.particle_data.hash_grid.update_grid()
    o.call("compute_nonpressure_force"); m.compute_nonpressure_force()
    o.call("init_iter_info"); m.init_iter_info()
    for it in range(3):
        o.call("update_iter_info"); m.update_iter_info()
        for f in ("pos_star", "vel_star"):
            a, b = util.eng_field(m, f), o.field(f)
            print(s, it, f, "maxabs", np.abs(a - b).max())
        o.call("predict_density"); m.predict_density()
        for f in ("adv_rho", "pressure", "d_vel_pre"):
            a, b = util.eng_field(m, f), o.field(f)
            e = np.abs(a - b)
            k = np.unravel_index(e.argmax(), e.shape)
            print(s, it, f, "max|b|", np.abs(b).max(), "maxabs", e.max(), "at", k, "a", a[k], "b", b[k])
    print("status", m.particle_data.hash_grid.status(), "nl max", None)
    o.call("update_pos"); m.update_pos()
