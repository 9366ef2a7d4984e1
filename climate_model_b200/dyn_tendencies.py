"""compute_tendencies (reference: dyn_tendencies.py:25-72): continuity -> momentum ->
temperature -> moisture through the factories, with the reference's timer keys."""
from .dyn_org_discretizations import TendencyFactory
from .io_read_namelist import B200

Tendencies = TendencyFactory(target=B200)


def compute_tendencies(GR, F):
    t = Tendencies.target
    GR.timer.start('cont')
    Tendencies.continuity(GR, GR.GRF[t], **F.get(Tendencies.fields_continuity, target=t))
    GR.timer.stop('cont')
    GR.timer.start('wind')
    Tendencies.momentum(GR.GRF[t], **F.get(Tendencies.fields_momentum, target=t))
    GR.timer.stop('wind')
    GR.timer.start('temp')
    Tendencies.temperature(GR.GRF[t], **F.get(Tendencies.fields_temperature, target=t))
    GR.timer.stop('temp')
    GR.timer.start('moist')
    if GR.i_moist_main_switch:
        Tendencies.moisture(GR.GRF[t], **F.get(Tendencies.fields_moisture, target=t))
    GR.timer.stop('moist')
