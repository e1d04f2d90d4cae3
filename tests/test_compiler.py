"""MJCF compiler: link counts, ordered joints, masses and collision filtering."""
import numpy as np
import pytest

from pybullet_gym_b200.mjcf import compiler as mj

EXPECT = {
    # xml: (Bullet links excl. base, parts, reduced bodies, total dofs, ground geoms, total mass)
    "inverted_pendulum.xml": (4, 4, 2, 2, 0, 15.49),
    "hopper.xml": (10, 10, 6, 6, 4, 15.82),
    "walker2d.xml": (16, 16, 9, 9, 7, 23.68),
    "half_cheetah.xml": (16, 16, 9, 9, 8, 38.19),
    "ant.xml": (20, 21, 9, 14, 13, 182.18),
    "humanoid_symmetric.xml": (29, 30, 18, 23, 17, 42.12),
}


@pytest.mark.parametrize("xml", sorted(EXPECT))
def test_structure(xml):
    links, parts, nb, nd, ground, mass = EXPECT[xml]
    bm, rm = mj.load(xml)
    assert len(bm.links) - 1 == links
    assert len(bm.part_names()) == parts
    assert rm.nb == nb and rm.nd == nd
    assert int(rm.geom_ground.sum()) == ground
    assert abs(sum(l.mass for l in bm.links) - mass) < 0.01
    assert abs(rm.mass.sum() - mass) < 0.01
    # subtrees are contiguous (bodies in DFS order) and parents precede children
    for b in range(rm.nb):
        assert rm.parent[b] < b
        assert rm.subtree_end[b] > b


def test_ordered_joints_follow_xml_body_order():
    bm, _ = mj.load("ant.xml")
    names = [bm.links[i].joint_name for i in bm.ordered_joints()]
    assert names == ["hip_1", "ankle_1", "hip_2", "ankle_2", "hip_3", "ankle_3", "hip_4", "ankle_4"]
    bm, _ = mj.load("hopper.xml")
    assert [bm.links[i].joint_name for i in bm.dof_links()][:3] == ["ignore1", "ignore2", "ignore3"]
    assert [bm.links[i].joint_name for i in bm.ordered_joints()] == ["thigh_joint", "leg_joint", "foot_joint"]


def test_limits_and_angle_units():
    bm, _ = mj.load("ant.xml")
    l = bm.links[bm.link_index("link0_%d" % 0) if False else [i for i, k in enumerate(bm.links) if k.joint_name == "ankle_1"][0]]
    assert np.isclose(l.lower, np.deg2rad(30)) and np.isclose(l.upper, np.deg2rad(100))
    bm, _ = mj.load("half_cheetah.xml")      # <compiler angle="radian">
    l = [k for k in bm.links if k.joint_name == "bthigh"][0]
    assert np.isclose(l.lower, -0.52) and np.isclose(l.upper, 1.05)
    bm, _ = mj.load("inverted_pendulum.xml")
    hinge = [k for k in bm.links if k.joint_name == "hinge"][0]
    assert hinge.lower > hinge.upper      # limited="false": pybullet reports (0, -1)


def test_collision_filter_or_rule():
    # conaffinity=0 robots have no self collision; humanoid collides with itself minus all ancestors
    for xml in ("ant.xml", "walker2d.xml", "half_cheetah.xml", "hopper.xml", "inverted_pendulum.xml"):
        assert len(mj.load(xml)[1].pair_a) == 0
    bm, rm = mj.load("humanoid_symmetric.xml")
    assert len(rm.pair_a) == 66
    torso_geoms = {g for g in range(len(rm.geom_body)) if rm.sub_names[rm.geom_link[g]] == "torso"}
    assert not (torso_geoms & set(rm.pair_a)) and not (torso_geoms & set(rm.pair_b))
    assert mj.load("inverted_pendulum.xml")[1].geom_ground.sum() == 0      # contype 0: contact-free


def test_reduction_preserves_mass_properties():
    """Merged bodies carry the same total mass, COM and inertia about the COM as their Bullet links."""
    for xml in EXPECT:
        bm, rm = mj.load(xml)
        R0, p0 = mj.link_world_frames(bm)
        m = sum(l.mass for l in bm.links)
        com = sum(l.mass * (p0[i] + R0[i] @ l.com) for i, l in enumerate(bm.links)) / m
        I = np.zeros((3, 3))
        for i, l in enumerate(bm.links):
            c = p0[i] + R0[i] @ l.com - com
            I += R0[i] @ np.diag(l.inertia) @ R0[i].T + l.mass * ((c @ c) * np.eye(3) - np.outer(c, c))
        # same quantities from the reduced tree at q = 0
        Rb, xb = [None] * rm.nb, [None] * rm.nb
        for b in range(rm.nb):
            q0 = mj.q_to_mat(rm.q0[b])
            if rm.parent[b] < 0:
                Rb[b] = q0
                A = rm.anchor_p[b]
            else:
                Rb[b] = Rb[rm.parent[b]] @ q0
                A = xb[rm.parent[b]] + Rb[rm.parent[b]] @ rm.anchor_p[b]
            xb[b] = A + Rb[b] @ rm.com_off[b]
        com2 = sum(rm.mass[b] * xb[b] for b in range(rm.nb)) / rm.mass.sum()
        I2 = np.zeros((3, 3))
        for b in range(rm.nb):
            ii = rm.inertia[b]
            Ib = np.array([[ii[0], ii[3], ii[4]], [ii[3], ii[1], ii[5]], [ii[4], ii[5], ii[2]]])
            c = xb[b] - com2
            I2 += Rb[b] @ Ib @ Rb[b].T + rm.mass[b] * ((c @ c) * np.eye(3) - np.outer(c, c))
        assert np.allclose(com, com2, atol=1e-12)
        assert np.allclose(I, I2, atol=1e-10)


def test_assets_are_canonical_emissions_of_reference_files():
    """If the reference checkout is present, the shipped assets compile to the same tables as the originals."""
    import os
    ref = "/root/reference/pybulletgym/envs/assets/mjcf"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    for xml in EXPECT:
        a, b = mj.reduce_model(mj.parse_mjcf(xml)), mj.reduce_model(mj.parse_mjcf(os.path.join(ref, xml)))
        for f in ("mass", "inertia", "anchor_p", "com_off", "axis", "jnt_lower", "jnt_upper", "geom_p0", "geom_p1", "geom_radius"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (xml, f)
