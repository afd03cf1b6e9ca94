// Land node and its surfaces (reference: wsimod/nodes/land.py, nodes/nutrient_pool.py).
//
// A surface's soil-water VQIP, its nutrient pools and its three outgoing flow VQIPs are held
// in registers while its inflow/process/outflow function lists run; every parameter plane is
// read exactly once where it is used.  Included by wsi_device.cuh.
#pragma once

namespace wsi {

// Test hook of the HOST build only (tests/hostsim, -DWSI_HOSTSIM_PROCESS_MASK): which entries of a surface's
// inflows / processes lists run, so that the reference's per-process known answers (tests/test_land.py:627-1069:
// one method of a GrowingSurface called on a prepared state) can be reproduced one process at a time through
// surface_phase.  In every CUDA compilation WSI_PROC(...) is the constant true and folds away.
#if !defined(__CUDACC__) && defined(WSI_HOSTSIM_PROCESS_MASK)
extern unsigned g_wsi_process_mask;
#define WSI_PROC(bit) (((g_wsi_process_mask >> (bit)) & 1u) != 0)
#else
#define WSI_PROC(bit) true
#endif
enum {
  WSI_PROC_CROP_COVER = 0, WSI_PROC_DRY_DEPOSITION, WSI_PROC_WET_DEPOSITION, WSI_PROC_SIMPLE_DEPOSITION, WSI_PROC_IHACRES,
  WSI_PROC_FLUSHING, WSI_PROC_FERTILISER, WSI_PROC_MANURE, WSI_PROC_SOIL_TEMPERATURE, WSI_PROC_TEMPERATURE_FACTOR,
  WSI_PROC_MOISTURE_FACTOR, WSI_PROC_SOIL_POOL, WSI_PROC_CROP_UPTAKE, WSI_PROC_EROSION, WSI_PROC_DENITRIFICATION,
  WSI_PROC_ADSORPTION
};

// vol + additive accumulation used for running_inflow_mb / running_outflow_mb
// (land.py:441-449): the non-additive lanes of those VQIPs never reach a comparison.
template <class C>
WSI_HD void mb_acc(const C& c, typename C::Vq& acc, const typename C::Vq& v) {
#pragma unroll
  for (int p = 0; p <= C::P; p++)
    if (p == 0 || bit(c.addm(), p)) acc.x[p] += v.x[p];
}

// NutrientStore.extract nutrient_pool.py:542-556
WSI_HD double pool_extract(double& store, double amount) {
  const double r = pymin(store, amount);
  store -= r;
  return r;
}

// NutrientPool state of one GrowingSurface, [N, P] pairs
struct Pools {
  double fast[2], humus[2], dis_inorg[2], dis_org[2], ads_inorg[2];
  double tdf, smdf;  // temperature / soil-moisture dependence factors
};

// temp_soil_process nutrient_pool.py:452-481
WSI_HD void temp_soil_process(const Pools& pl, const double par0, const double par1, double* ext,
                              double* rcv, double* amount) {
  amount[0] = par0 * pl.tdf * pl.smdf * ext[0];
  amount[1] = par1 * pl.tdf * pl.smdf * ext[1];
  amount[0] = pool_extract(ext[0], amount[0]);
  amount[1] = pool_extract(ext[1], amount[1]);
  rcv[0] += amount[0];
  rcv[1] += amount[1];
}

// GrowingSurface.quick_interp land.py:1495-1512: bisect_left + Python negative-index wrap.
// xp/yp are table planes (stride `st` doubles apart)
WSI_HD double quick_interp(double x, const double* xp, const double* yp, size_t st, int n) {
  int i = 0;
  while (i < n && ldp(xp + (size_t)i * st) < x) i++;
  int il = i - 1;
  if (il < 0) il += n;
  if (i >= n) i = n - 1;  // the reference raises IndexError here
  const double x_left = ldp(xp + (size_t)il * st), x_right = ldp(xp + (size_t)i * st);
  const double y_left = ldp(yp + (size_t)il * st), y_right = ldp(yp + (size_t)i * st);
  return y_left + (y_right - y_left) * (x - x_left) / (x_right - x_left);
}

// GrowingSurface.adjust_vqip_to_liquid land.py:1574-1603
template <class C>
WSI_HD void adjust_to_liquid(const C& c, typename C::Vq& v, double depN, double depP, double inN,
                             double inP) {
  if (c.cfg.no3() < 0) return;
  if (depN > 0) {
    vmul<C::P>(v, c.cfg.no3(), inN / depN);
    vmul<C::P>(v, c.cfg.nh4(), inN / depN);
    vmul<C::P>(v, c.cfg.orgn(), inN / depN);
  }
  if (depP > 0) {
    vmul<C::P>(v, c.cfg.po4(), inP / depP);
    vmul<C::P>(v, c.cfg.orgp(), inP / depP);
  }
}

// One surface of a Land node: Surface.run land.py:425-449 with the function lists of
// Surface / ImperviousSurface / PerviousSurface / GrowingSurface (land.py:375-382, 619-621,
// 830-841, 1275-1291).  Runs as its own thread (wsi_surface_kernel: one thread per surface, a
// block is homogeneous in surface type) or inline in the per-replica step.  What the parent
// Land needs afterwards goes through the surface's scratch planes: the flows of `route`
// (land.py:1046-1058) or the ponded runoff of `push_to_sewers` (land.py:699-721), and its share
// of running_inflow_mb / running_outflow_mb / ds.  TYPE >= 0 fixes the surface class at compile
// time (-1: read it from the type plane).
template <int TYPE, class C>
WSI_HD void surface_phase(C& c, int u) {
  using Vq = typename C::Vq;
  constexpr int P = C::P;
  constexpr int V = C::V;
  const size_t ss = c.s_surf();
  const Obj S = c.surf_obj(u);
  const int type = TYPE >= 0 ? TYPE : (int)ldp(S(WSI_SURF_TYPE));
  const int L = (int)ldp(S(WSI_SURF_LAND));
  const Obj Lo = c.land_obj(L);
  const int fid = (int)*Lo(WSI_LAND_FORCING_ID);
  const double precip = c.forcing(fid, 0), et0 = c.forcing(fid, 1), temperature = c.forcing(fid, 2);
  const bool want_ds = (c.run_flags & WSI_RUN_MASS_BALANCE) != 0;
  Vq mb_in = vq_zero<P>(), mb_out = vq_zero<P>();
  const double area = ldp(S(WSI_SURF_AREA));
  double* stp = S(WSI_SURF_STORAGE(V));
  Vq st = c.ldv(stp, ss);
  const auto& cfg = c.cfg;

  // nitrification hand-over from last step's decay (land.py:427-439)
  if (cfg.no2() >= 0) {
    const double dn2 = decay_total_of(S, WSI_SURF_N_DECAY, cfg.no2());
    vadd<P>(st, cfg.no3(), dn2);
    vadd<P>(mb_in, cfg.no3(), dn2);
    const double dnh = decay_total_of(S, WSI_SURF_N_DECAY, cfg.nh4());
    vadd<P>(st, cfg.no2(), dnh);
    vadd<P>(mb_in, cfg.no2(), dnh);
  }

  const bool growing = type == WSI_ST_GROWINGSURFACE;
  const bool nutrients = growing && cfg.no3() >= 0;
  Pools pl;
  if (growing) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
      pl.fast[k] = *S(WSI_SURF_POOL_FAST + k);
      pl.humus[k] = *S(WSI_SURF_POOL_HUMUS + k);
      pl.dis_inorg[k] = *S(WSI_SURF_POOL_DIS_INORG + k);
      pl.dis_org[k] = *S(WSI_SURF_POOL_DIS_ORG + k);
      pl.ads_inorg[k] = *S(WSI_SURF_POOL_ADS_INORG + k);
    }
    pl.tdf = 0.0;
    pl.smdf = 0.0;
  }
  double et0_coef = *S(WSI_SURF_ET0_COEFFICIENT);
  double days_after_sow = 0.0, crop_cover = 0.0, ground_cover = 0.0, crop_factor_out = 0.0;

  // ---- inflows --------------------------------------------------------------------------
  if (growing && WSI_PROC(WSI_PROC_CROP_COVER)) {  // calc_crop_cover land.py:1514-1572
    const double doy = (double)c.D().doy_adj[c.t];
    const double dayofyear = (double)c.D().dayofyear[c.t];
    days_after_sow = *S(WSI_SURF_DAYS_AFTER_SOW);
    crop_cover = *S(WSI_SURF_CROP_COVER);
    ground_cover = *S(WSI_SURF_GROUND_COVER);
    if (days_after_sow < 0) {
      if (dayofyear == ldp(S(WSI_SURF_SOWING_DAY))) days_after_sow = 0.0;
    } else {
      if (dayofyear == ldp(S(WSI_SURF_HARVEST_DAY))) {
        days_after_sow = -1.0;
        crop_cover = 0.0;
        ground_cover = 0.0;
      } else {
        days_after_sow += 1.0;
      }
    }
    const double crop_factor = quick_interp(doy, S(WSI_SURF_STAGE_DATES), S(WSI_SURF_STAGE_VALS), ss,
                                            (int)ldp(S(WSI_SURF_N_STAGE)));
    if (days_after_sow > 0) {
      crop_cover = quick_interp(doy, S(WSI_SURF_HS_CALENDAR), S(WSI_SURF_CROP_COVER_STAGES), ss, 5);
      ground_cover = quick_interp(doy, S(WSI_SURF_HS_CALENDAR), S(WSI_SURF_GROUND_COVER_STAGES), ss, 5);
    }
    const double fc_m = ldp(S(WSI_SURF_FIELD_CAPACITY_M));
    const double rzd = pymax(fc_m - st.x[0] / area, 0.0);
    double stress;
    if (rzd < ldp(S(WSI_SURF_READILY_AVAILABLE_WATER))) {
      stress = 1.0;
    } else {
      const double taw = ldp(S(WSI_SURF_TOTAL_AVAILABLE_WATER));
      stress = pymax(0.0, (taw - rzd) / ((1 - ldp(S(WSI_SURF_ET_DEPLETION_FACTOR))) * taw));
    }
    et0_coef = stress * crop_factor;
    crop_factor_out = crop_factor;
  }

  const int dep_id = (int)ldp(S(WSI_SURF_DEP_ID));
  if (ldp(S(WSI_SURF_HAS_DEP)) != 0.0) {
    // atmospheric_deposition / precipitation_deposition land.py:528-581 (dry, then wet)
#pragma unroll
    for (int wet = 0; wet < 2; wet++) {
      if (!WSI_PROC(wet ? WSI_PROC_WET_DEPOSITION : WSI_PROC_DRY_DEPOSITION)) continue;
      const int b = wet ? WSI_DEP_NHX_WET : WSI_DEP_NHX_DRY;
      const double nhx = c.dep(dep_id, b) * area;
      const double noy = c.dep(dep_id, b + 1) * area;
      const double srp = c.dep(dep_id, b + 2) * area;
      Vq v = vq_zero<P>();
      vset<P>(v, cfg.nh4(), nhx);
      vset<P>(v, cfg.no3(), noy);
      vset<P>(v, cfg.po4(), srp);
      if (growing) {  // dry/wet_deposition_to_tank land.py:2267-2311
        const double depN = noy + nhx, depP = srp;
        double inN, inP;
        if (!wet) {  // allocate_dry_deposition nutrient_pool.py:226-248
          const double f_di = ldp(S(WSI_SURF_FRAC_DRY_N_TO_DIS_INORG));
          pl.fast[0] += depN * ldp(S(WSI_SURF_FRAC_DRY_N_TO_FAST));
          pl.dis_inorg[0] += depN * f_di;
          pl.ads_inorg[1] += depP;
          inN = depN * f_di;
          inP = 0.0;
        } else {  // allocate_wet_deposition nutrient_pool.py:250-265
          pl.dis_inorg[0] += depN;
          pl.dis_inorg[1] += depP;
          inN = depN;
          inP = depP;
        }
        adjust_to_liquid(c, v, depN, depP, inN, inP);
      }
      st = c.sum(st, v);
      mb_acc(c, mb_in, v);
    }
  }
  if (ldp(S(WSI_SURF_HAS_LOAD)) != 0.0 && WSI_PROC(WSI_PROC_SIMPLE_DEPOSITION)) {  // simple_deposition land.py:505-526
    Vq pol;
    pol.x[0] = 0.0;
    const double* lp = S(WSI_SURF_LOAD(V));
#pragma unroll
    for (int p = 1; p <= P; p++) pol.x[p] = bit(c.addm(), p) ? ldp(lp + (size_t)p * ss) * area : ldp(lp + (size_t)p * ss);
    st = c.sum(st, pol);
    mb_acc(c, mb_in, pol);
  }

  Vq ie = vq_zero<P>(), sf = vq_zero<P>(), pc = vq_zero<P>();  // infiltration_excess, subsurface_flow, percolation
  double out_evap = 0.0, out_precip = 0.0, out_recharge = 0.0;

  if (type == WSI_ST_IMPERVIOUSSURFACE) {
    // precipitation_evaporation land.py:645-697
    const double evaporation_depth = et0 * ldp(S(WSI_SURF_ET0_TO_E));
    double total_evaporation;
    if (precip < evaporation_depth) {
      double e = (evaporation_depth - precip) * area;
      e = pymin(e, st.x[0]);  // Tank.evaporate tanks.py:303-317
      st.x[0] -= e;
      total_evaporation = e + precip * area;
    } else {
      Vq np = vq_volume<P>((precip - evaporation_depth) * area);
      vset<P>(np, cfg.temp(), temperature);
      st = c.sum(st, np);
      total_evaporation = evaporation_depth * area;
    }
    out_evap = total_evaporation;
    out_precip = precip * area;
    mb_in.x[0] += out_precip;
    mb_out.x[0] += out_evap;
    // push_to_sewers land.py:699-721, first half: pull_ponded.  The push to the Sewers and
    // forcing the reply back into this tank happen in land_phase (the sewer is shared state).
    const double ponded = pymax(st.x[0] - ldp(S(WSI_SURF_CAPACITY)), 0.0);
    ie = vq_pull(c, st, ponded);
  } else if (type >= WSI_ST_PERVIOUSSURFACE) {
    // ihacres land.py:913-1044
    const double depth = ldp(S(WSI_SURF_DEPTH));
    const double capacity = ldp(S(WSI_SURF_CAPACITY));
    const double fc_m = ldp(S(WSI_SURF_FIELD_CAPACITY_M));
    const double wp_m = ldp(S(WSI_SURF_WILTING_POINT_M));
    if (WSI_PROC(WSI_PROC_IHACRES)) {
    double evaporation_depth = et0 * et0_coef;  // get_climate land.py:903-911
    double precip_s = precip;
    if (c.D().features & 2) {  // VariableAreaSurface.get_climate_ land.py:2449-2461 (1 for every other class)
      const double scale = ldp(S(WSI_SURF_CLIMATE_SCALE));
      precip_s *= scale;
      evaporation_depth *= scale;
    }
    const double inf_cap = ldp(S(WSI_SURF_INFILTRATION_CAPACITY));
    const double infiltrated = pymin(precip_s, inf_cap);
    double infiltration_excess = pymax(precip_s - infiltrated, 0.0);
    const double cmd = tank_excess_vol(capacity, st.x[0], false, 0.0) / area;  // get_cmd land.py:886
    double evaporation = evaporation_depth * pymin(1.0, exp(2 * (1 - cmd / (depth - wp_m))));
    const double outflow =
        infiltrated * (1 - pymin(1.0, pow(cmd / (depth - fc_m), ldp(S(WSI_SURF_IHACRES_P)))));
    evaporation = pymin(evaporation, precip_s + st.x[0] / area);
    const double surf_coef = ldp(S(WSI_SURF_SURFACE_COEFFICIENT));
    const double surface = outflow * surf_coef * area;
    const double percolation = outflow * (1 - surf_coef) * ldp(S(WSI_SURF_PERCOLATION_COEFFICIENT)) * area;
    const double subsurface_flow = outflow * (1 - surf_coef) * ldp(S(WSI_SURF_SUBSURFACE_COEFFICIENT)) * area;
    const double tank_recharge = (infiltrated - evaporation - outflow) * area;
    infiltration_excess *= area;
    infiltration_excess += surface;
    evaporation *= area;
    const double precipitation = precip_s * area;
    const double total = tank_recharge + subsurface_flow + percolation;
    if (total > 0) {
      Vq tw = vq_volume<P>(total);
      vset<P>(tw, cfg.temp(), temperature);
      st = c.sum(st, tw);
      // pull_storage: Tank version (tanks.py:207-239) or GrowingSurface override (land.py:1433-1493)
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const double want = k == 0 ? subsurface_flow : percolation;
        Vq reply = vq_zero<P>();
        if (st.x[0] != 0) {
          reply = c.vchange(st, pymin(want, st.x[0]));
          if (nutrients) {
            const double prop = reply.x[0] / st.x[0];
            // extract_dissolved nutrient_pool.py:330-356
            const double diN = pool_extract(pl.dis_inorg[0], pl.dis_inorg[0] * prop);
            const double diP = pool_extract(pl.dis_inorg[1], pl.dis_inorg[1] * prop);
            const double doN = pool_extract(pl.dis_org[0], pl.dis_org[0] * prop);
            const double doP = pool_extract(pl.dis_org[1], pl.dis_org[1] * prop);
            const double sno3 = vget<P>(st, cfg.no3()), snh4 = vget<P>(st, cfg.nh4()),
                         sno2 = vget<P>(st, cfg.no2());
            const double den = sno3 + snh4 + sno2;
            if (den == 0) c.flag((int)*Lo(WSI_LAND_NODE), WSI_FLAG_NAN);
            vset<P>(reply, cfg.no3(), diN * sno3 / den);
            vset<P>(reply, cfg.nh4(), diN * snh4 / den);
            vset<P>(reply, cfg.no2(), diN * sno2 / den);
            vset<P>(reply, cfg.po4(), diP);
            vset<P>(reply, cfg.orgp(), doP);
            vset<P>(reply, cfg.orgn(), doN);
          }
          st = c.extract(st, reply);
        }
        if (k == 0)
          sf = reply;
        else
          pc = reply;
      }
    } else {
      const double e = pymin(-total, st.x[0]);  // Tank.evaporate
      st.x[0] -= e;
    }
    ie = vq_volume<P>(infiltration_excess);
    vset<P>(ie, cfg.temp(), temperature);
    out_evap = evaporation;
    out_precip = precipitation;
    out_recharge = tank_recharge;
    mb_in.x[0] += precipitation;
    mb_out.x[0] += evaporation;
    }

    if (nutrients) {
      // effective_precipitation_flushing land.py:1605-1647
      if (WSI_PROC(WSI_PROC_FLUSHING)) {
        const int a = cfg.nh4(), b = cfg.no2(), n3 = cfg.no3();
        double N = vget<P>(sf, a) + vget<P>(sf, b) + vget<P>(sf, n3) + vget<P>(pc, a) + vget<P>(pc, b) +
                   vget<P>(pc, n3) + vget<P>(ie, a) + vget<P>(ie, b) + vget<P>(ie, n3);
        double Pp = vget<P>(sf, cfg.po4()) + vget<P>(pc, cfg.po4()) + vget<P>(ie, cfg.po4());
        (void)pool_extract(pl.dis_inorg[0], N);
        (void)pool_extract(pl.dis_inorg[1], Pp);
        N = vget<P>(sf, cfg.orgn()) + vget<P>(pc, cfg.orgn()) + vget<P>(ie, cfg.orgn());
        Pp = vget<P>(sf, cfg.orgp()) + vget<P>(pc, cfg.orgp()) + vget<P>(ie, cfg.orgp());
        (void)pool_extract(pl.dis_org[0], N);
        (void)pool_extract(pl.dis_org[1], Pp);
      }
      // fertiliser land.py:1649-1679, manure land.py:1681-1710
#pragma unroll
      for (int manure = 0; manure < 2; manure++) {
        if (!WSI_PROC(manure ? WSI_PROC_MANURE : WSI_PROC_FERTILISER)) continue;
        const int b = manure ? WSI_DEP_NHX_MANURE : WSI_DEP_NHX_FERTILISER;
        const double nhx = c.dep(dep_id, b) * area;
        const double noy = c.dep(dep_id, b + 1) * area;
        const double srp = c.dep(dep_id, b + 2) * area;
        Vq v = vq_zero<P>();
        vset<P>(v, cfg.nh4(), nhx);
        vset<P>(v, cfg.no3(), noy);
        vset<P>(v, cfg.po4(), srp);
        const double depN = noy + nhx, depP = srp;
        double inN, inP;
        if (!manure) {  // allocate_fertiliser nutrient_pool.py:314-328
          pl.dis_inorg[0] += depN;
          pl.dis_inorg[1] += depP;
          inN = depN;
          inP = depP;
        } else {  // allocate_manure nutrient_pool.py:267-290
          const double fN = ldp(S(WSI_SURF_FRAC_MANURE_TO_DIS_INORG)), fP = ldp(S(WSI_SURF_FRAC_MANURE_TO_DIS_INORG + 1));
          pl.dis_inorg[0] += depN * fN;
          pl.dis_inorg[1] += depP * fP;
          pl.fast[0] += depN * ldp(S(WSI_SURF_FRAC_MANURE_TO_FAST));
          pl.fast[1] += depP * ldp(S(WSI_SURF_FRAC_MANURE_TO_FAST + 1));
          inN = depN * fN;
          inP = depP * fP;
        }
        adjust_to_liquid(c, v, depN, depP, inN, inP);
        st = c.sum(st, v);
        mb_acc(c, mb_in, v);
      }
    }

    // ---- processes ----------------------------------------------------------------------
    if (cfg.temp() >= 0 && WSI_PROC(WSI_PROC_SOIL_TEMPERATURE)) {  // calculate_soil_temperature land.py:1060-1079
      const double w_prev = ldp(S(WSI_SURF_SOIL_TEMP_W_PREV)), w_air = ldp(S(WSI_SURF_SOIL_TEMP_W_AIR)),
                   w_deep = ldp(S(WSI_SURF_SOIL_TEMP_W_DEEP));
      const double au = vget<P>(st, cfg.temp()) * w_prev;
      const double air = temperature * w_air;
      const double total_weight = w_air + w_deep + w_prev;
      vset<P>(st, cfg.temp(), (au + air + ldp(S(WSI_SURF_SOIL_TEMP_DEEP)) * w_deep) / total_weight);
    }
    if (nutrients) {
      const double rooting_depth = ldp(S(WSI_SURF_ROOTING_DEPTH));
      const double soilT = vget<P>(st, cfg.temp());
      // calc_temperature_dependence_factor land.py:1813-1831
      if (!WSI_PROC(WSI_PROC_TEMPERATURE_FACTOR))
        pl.tdf = 0.0;
      else if (soilT > 5)
        pl.tdf = pow(2.0, (soilT - 20) / 10);
      else if (soilT > 0)
        pl.tdf = soilT / 5;
      else
        pl.tdf = 0.0;
      // calc_soil_moisture_dependence_factor land.py:1833-1856
      if (WSI_PROC(WSI_PROC_MOISTURE_FACTOR)) {
        const double smc = st.x[0] / area;
        const double satact = ldp(S(WSI_SURF_SATACT));
        if (smc >= fc_m) {
          pl.smdf = satact;
        } else if (smc <= wp_m) {
          pl.smdf = 0.0;
        } else {
          const double thetapow = ldp(S(WSI_SURF_THETAPOW));
          double fc_comp = pow((fc_m - smc) / (ldp(S(WSI_SURF_THETAUPP)) * rooting_depth), thetapow);
          fc_comp = (1 - satact) * fc_comp + satact;
          const double wp_comp = pow((smc - wp_m) / (ldp(S(WSI_SURF_THETALOW)) * rooting_depth), thetapow);
          pl.smdf = pymin(pymin(1.0, wp_comp), fc_comp);
        }
      }
      // soil_pool_transformation land.py:1744-1811 + nutrient_pool.py:396-450
      if (WSI_PROC(WSI_PROC_SOIL_POOL)) {
        const double sno3 = vget<P>(st, cfg.no3()), snh4 = vget<P>(st, cfg.nh4());
        if (sno3 + snh4 == 0) c.flag((int)*Lo(WSI_LAND_NODE), WSI_FLAG_NAN);
        const double nitrate_proportion = sno3 / (sno3 + snh4);
        double inc_di[2] = {0.0, 0.0}, inc_do[2] = {0.0, 0.0}, amount[2];
        temp_soil_process(pl, ldp(S(WSI_SURF_DEGRHPAR)), ldp(S(WSI_SURF_DEGRHPAR + 1)), pl.humus, pl.fast, amount);
        temp_soil_process(pl, ldp(S(WSI_SURF_DISHPAR)), ldp(S(WSI_SURF_DISHPAR + 1)), pl.humus, pl.dis_org, amount);
        inc_do[0] = inc_do[0] + amount[0];
        inc_do[1] = inc_do[1] + amount[1];
        temp_soil_process(pl, ldp(S(WSI_SURF_MINFPAR)), ldp(S(WSI_SURF_MINFPAR + 1)), pl.fast, pl.dis_inorg, amount);
        inc_di[0] = inc_di[0] + amount[0];
        inc_di[1] = inc_di[1] + amount[1];
        temp_soil_process(pl, ldp(S(WSI_SURF_DISFPAR)), ldp(S(WSI_SURF_DISFPAR + 1)), pl.fast, pl.dis_org, amount);
        inc_do[0] = inc_do[0] + amount[0];
        inc_do[1] = inc_do[1] + amount[1];
        temp_soil_process(pl, ldp(S(WSI_SURF_IMMOBDPAR)), ldp(S(WSI_SURF_IMMOBDPAR + 1)), pl.dis_inorg, pl.fast, amount);
        inc_di[0] = inc_di[0] - amount[0];
        inc_di[1] = inc_di[1] - amount[1];
        Vq in_ = vq_zero<P>(), out_ = vq_zero<P>();
        if (inc_di[0] > 0) {
          vset<P>(in_, cfg.no3(), inc_di[0] * nitrate_proportion);
          vset<P>(in_, cfg.nh4(), inc_di[0] * (1 - nitrate_proportion));
        } else {
          vset<P>(out_, cfg.no3(), -inc_di[0] * nitrate_proportion);
          vset<P>(out_, cfg.nh4(), -inc_di[0] * (1 - nitrate_proportion));
        }
        if (inc_do[0] > 0)
          vset<P>(in_, cfg.orgn(), inc_do[0]);
        else
          vset<P>(out_, cfg.orgn(), -inc_do[0]);
        if (inc_di[1] > 0)
          vset<P>(in_, cfg.po4(), inc_di[1]);
        else
          vset<P>(out_, cfg.po4(), -inc_di[1]);
        if (inc_do[1] > 0)
          vset<P>(in_, cfg.orgp(), inc_do[1]);
        else
          vset<P>(out_, cfg.orgp(), -inc_do[1]);
        st = c.sum(st, in_);
        vq_pull_pollutants(c, st, out_);  // clips out_ in place (tanks.py:261-262)
        // land.py:1806-1808 "nutrient pool not tracking soil tank": pull_pollutants hands its
        // (clipped) argument back, so the reference compares the request with itself -- the
        // advisory can only fire on a NaN-free difference, i.e. never; kept as the status bit
        if (!vq_compare(c, out_, out_)) c.flag((int)*Lo(WSI_LAND_NODE), WSI_FLAG_POOL_MISMATCH);
        mb_acc(c, mb_in, in_);
        mb_acc(c, mb_out, out_);
      }
      // calc_crop_uptake land.py:1858-1926
      if (days_after_sow > 0 && WSI_PROC(WSI_PROC_CROP_UPTAKE)) {
        double das = days_after_sow;
        double temp_func = 1.0;
        if (ldp(S(WSI_SURF_AUTUMN_SOW)) != 0.0) {
          temp_func = pymax(0.0, pymin(1.0, (vget<P>(st, cfg.temp()) - 5) / 20));
          das -= 25;
        }
        const double u1 = ldp(S(WSI_SURF_UPTAKE1)), u2 = ldp(S(WSI_SURF_UPTAKE2)), u3 = ldp(S(WSI_SURF_UPTAKE3));
        const double uptake_par = (u1 - u2) * exp(-u3 * das);
        double N_common = 0.0;
        if (uptake_par + u2 > 0) N_common = u1 * u2 * u3 * uptake_par / ((u2 + uptake_par) * (u2 + uptake_par));
        N_common *= temp_func * 1e-3 * area;  // G_M2_TO_KG_M2
        const double P_common = N_common * ldp(S(WSI_SURF_UPTAKE_PNRATIO));
        const double frac = pymax(0.0, st.x[0] - wp_m * area) / st.x[0];
        const double N_max = frac * pl.dis_inorg[0];
        const double P_max = frac * pl.dis_inorg[1];
        const double gotP = pool_extract(pl.dis_inorg[1], pymin(P_common, P_max));
        const double gotN = pool_extract(pl.dis_inorg[0], pymin(N_common, N_max));
        Vq out_ = vq_zero<P>();
        vset<P>(out_, cfg.no3(), gotN);
        vset<P>(out_, cfg.po4(), gotP);
        vq_pull_pollutants(c, st, out_);
        mb_acc(c, mb_out, out_);
      }
      // erosion land.py:1928-2094
      if (WSI_PROC(WSI_PROC_EROSION)) {
        const double precipitation_depth = precip * 1e3;  // M_TO_MM
        double mobilised_rain = 0.0, mobilised_flow = 0.0;
        if (precipitation_depth > 5) {
          const double dayofyear = (double)c.D().dayofyear[c.t];
          double rainfall_energy =
              8.95 + 8.44 * log10(precipitation_depth *
                                  (0.257 + sin(2 * 3.141592653589793 * ((dayofyear - 70) / 365)) * 0.09) * 2);
          rainfall_energy *= precipitation_depth;
          mobilised_rain = rainfall_energy * (1 - crop_cover) * ldp(S(WSI_SURF_ERODIBILITY));
        }
        if (ie.x[0] > 0) {
          mobilised_flow = pow(ie.x[0] / area * 1e3 * 365, ldp(S(WSI_SURF_SREROEXP)));
          mobilised_flow *= (1 - ground_cover) * (1 / (0.5 * ldp(S(WSI_SURF_COHESION)))) *
                            sin(ldp(S(WSI_SURF_SLOPE)) / 100) / 365;
        }
        const double total_flows = ie.x[0] + sf.x[0] + pc.x[0];
        const double erodingflow = total_flows / area * 1e3;
        const double transportfactor = pymin(1.0, pow(erodingflow / 4, 1.3));
        double erodedsed = 1000 * (mobilised_flow + mobilised_rain) * transportfactor;
        double enrichment = 0.0;
        bool active = true;
        if (erodingflow > 4)
          enrichment = 1.5;
        else if (erodingflow > 0)
          enrichment = 4 - (4 - 1.5) * erodingflow / 4;
        else
          active = false;  // early return land.py:1996-1997
        if (active) {
          const double erodableP = (pl.ads_inorg[1] + pl.humus[1]) / area * 1e6;  // KG_M2_TO_KG_KM2
          const double bulk_density = ldp(S(WSI_SURF_BULK_DENSITY));
          double erodedP = erodedsed * (erodableP / (rooting_depth * 1e-3 * bulk_density * 1e9)) * enrichment;
          erodedP *= area * 1e-6;
          erodedsed *= area * 1e-6;
          const double srfilt = ldp(S(WSI_SURF_SRFILT)), macrofilt = ldp(S(WSI_SURF_MACROFILT));
          const double surface_erodedP = srfilt * ie.x[0] / total_flows * erodedP;
          const double surface_erodedsed = srfilt * ie.x[0] / total_flows * erodedsed;
          const double subsurface_erodedP = macrofilt * sf.x[0] / total_flows * erodedP;
          const double subsurface_erodedsed = macrofilt * sf.x[0] / total_flows * erodedsed;
          const double percolation_erodedP = macrofilt * pc.x[0] / total_flows * erodedP;
          const double percolation_erodedsed = macrofilt * pc.x[0] / total_flows * erodedsed;
          const double eff_erodedP = percolation_erodedP + surface_erodedP + subsurface_erodedP;
          if (eff_erodedP > 0) {
            // erode_P nutrient_pool.py:367-394
            const double fraction_adsorbed = pl.ads_inorg[1] / (pl.ads_inorg[1] + pl.humus[1]);
            const double inorg_removed = pool_extract(pl.ads_inorg[1], eff_erodedP * fraction_adsorbed);
            const double org_removed = pool_extract(pl.humus[1], eff_erodedP * (1 - fraction_adsorbed));
            vadd<P>(ie, cfg.orgp(), surface_erodedP * org_removed / eff_erodedP);
            vadd<P>(sf, cfg.orgp(), subsurface_erodedP * org_removed / eff_erodedP);
            vadd<P>(pc, cfg.orgp(), percolation_erodedP * org_removed / eff_erodedP);
            vadd<P>(ie, cfg.po4(), surface_erodedP * inorg_removed / eff_erodedP);
            vadd<P>(sf, cfg.po4(), subsurface_erodedP * inorg_removed / eff_erodedP);
            vadd<P>(pc, cfg.po4(), percolation_erodedP * inorg_removed / eff_erodedP);
            vadd<P>(mb_in, cfg.po4(), inorg_removed);
            vadd<P>(mb_in, cfg.orgp(), org_removed);
          }
          if (cfg.solids() >= 0) {
            vadd<P>(ie, cfg.solids(), surface_erodedsed);
            vadd<P>(sf, cfg.solids(), subsurface_erodedsed);
            vadd<P>(pc, cfg.solids(), percolation_erodedsed);
            vadd<P>(mb_in, cfg.solids(), surface_erodedsed + subsurface_erodedsed + percolation_erodedsed);
          }
        }
      }
      // denitrification land.py:2096-2152
      if (WSI_PROC(WSI_PROC_DENITRIFICATION)) {
        const double smc = st.x[0] / area;
        double dep_f = -1.0;
        const double limpar = ldp(S(WSI_SURF_LIMPAR));
        if (smc > fc_m)
          dep_f = 1.0;
        else if (smc / fc_m > limpar)
          dep_f = pow(((smc / fc_m) - limpar) / (1 - limpar), ldp(S(WSI_SURF_EXPPAR)));
        if (dep_f >= 0) {
          double din_conc = pl.dis_inorg[0] / st.x[0];
          din_conc *= 1e3;  // KG_M3_TO_MG_L
          const double hs = din_conc / (din_conc + ldp(S(WSI_SURF_HSATINS)));
          const double denitrified = pl.dis_inorg[0] * hs * dep_f * pl.tdf * ldp(S(WSI_SURF_DENPAR));
          const double got = pool_extract(pl.dis_inorg[0], denitrified);
          Vq out_ = vq_zero<P>();
          vset<P>(out_, cfg.no3(), got);
          vq_pull_pollutants(c, st, out_);
          mb_acc(c, mb_out, out_);
        }
      }
      // adsorption land.py:2154-2265 (Freundlich isotherm, Newton-Raphson)
      if (WSI_PROC(WSI_PROC_ADSORPTION)) {
        const double bulk_density = ldp(S(WSI_SURF_BULK_DENSITY));
        double ad_de_P_pool = pl.ads_inorg[1] + pl.dis_inorg[1];
        ad_de_P_pool /= area * 1e-6;
        if (ad_de_P_pool != 0) {
          const double kfr = ldp(S(WSI_SURF_KFR)), nfr = ldp(S(WSI_SURF_NFR));
          const double smc_mm = st.x[0] / area * 1e3;
          const double conc_sol = pl.ads_inorg[1] * 1e6 / (bulk_density * rooting_depth * area);
          const double coeff = kfr * bulk_density * rooting_depth;
          double ad_P_equi_conc;
          if (conc_sol <= 0) {
            const double xn_1 = ad_de_P_pool / (smc_mm + coeff);
            ad_P_equi_conc = kfr * xn_1;
          } else {
            const double limit = ldp(S(WSI_SURF_ADSORPTION_NR_LIMIT));
            const int maxiter = (int)ldp(S(WSI_SURF_ADSORPTION_NR_MAXITER));
            const double x0 = exp((log(conc_sol) - log(kfr)) / nfr);
            double fxn = x0 * smc_mm + coeff * pow(x0, nfr) - ad_de_P_pool;
            double xn = x0, xn_1 = x0;
            int j = 0;
            while (fabs(fxn) > limit && j < maxiter) {
              fxn = xn * smc_mm + coeff * pow(xn, nfr) - ad_de_P_pool;
              const double fprimxn = smc_mm + nfr * coeff * pow(xn, nfr - 1);
              const double dx = fxn / fprimxn;
              if (fabs(dx) < 0.000001 * xn) break;
              xn_1 = xn - dx;
              if (xn_1 <= 0) xn_1 = 1e-10;
              xn = xn_1;
              j++;
            }
            ad_P_equi_conc = kfr * pow(xn_1, nfr);
          }
          if (fabs(ad_P_equi_conc - conc_sol) > 1e-6) {
            const double adsdes = (ad_P_equi_conc - conc_sol) * (1 - exp(-ldp(S(WSI_SURF_KADSDES))));
            double req = adsdes * bulk_density * rooting_depth * (area * 1e-6);
            if (req > 0) {
              const double adsorbed = pool_extract(pl.dis_inorg[1], req);
              pl.ads_inorg[1] += adsorbed;
              Vq out_ = vq_zero<P>();
              vset<P>(out_, cfg.po4(), adsorbed);
              vq_pull_pollutants(c, st, out_);
              mb_acc(c, mb_out, out_);
            } else {
              req = -req;
              const double desorbed = pool_extract(pl.ads_inorg[1], req);
              pl.dis_inorg[1] += desorbed;
              Vq in_ = vq_zero<P>();
              vset<P>(in_, cfg.po4(), desorbed);
              st = c.sum(st, in_);
              mb_acc(c, mb_in, in_);
            }
          }
        }
      }
    }
  }
  // ---- write back ---------------------------------------------------------------------------
  if (type != WSI_ST_SURFACE) {  // flows for the parent's tanks / sewers
    c.stv(S(WSI_SURF_FLOW_A(V)), ss, ie);
    if (type >= WSI_ST_PERVIOUSSURFACE) {
      c.stv(S(WSI_SURF_FLOW_B(V)), ss, sf);
      c.stv(S(WSI_SURF_FLOW_C(V)), ss, pc);
    }
  }
  if (want_ds) {
    // storage_ is what is still in memory (or what was, before this step's garden irrigation arrived),
    // except for the decayed pollutants (pre_decay)
    const bool irrigated = growing && *S(WSI_SURF_IRRIGATED_AT) == (double)c.t;
    Vq old = c.ldv_add(irrigated ? S(WSI_SURF_PRE_IRRIGATION(V)) : stp, ss);
    const int n = (int)*S(WSI_SURF_N_DECAY);
    for (int i = 0; i < WSI_MAX_DECAY; i++) {
      if (i >= n) break;
      vset<P>(old, (int)*S(WSI_SURF_DECAY_POL + i), *S(WSI_SURF_PRE_DECAY + i));
    }
    // an IrrigationSurface is watered AFTER it has run (Land.apply_irrigation, land.py:192-198, 2347-2381):
    // land_irrigate re-forms this step's ds from the storage_ kept here
    if (growing && (c.D().features & 2) && *S(WSI_SURF_IRRIGATION_COEFFICIENT) >= 0.0)
      c.stv_add(S(WSI_SURF_PRE_IRRIGATION(V)), ss, old);
    if (type == WSI_ST_IMPERVIOUSSURFACE) {
      // the tank still receives the sewers' reply: land_phase forms ds from storage_
      c.stv_add(S(WSI_SURF_MB_DS(V)), ss, old);
    } else {
      // decay_ds tanks.py:449-457: (storage - storage_) + total_decayed
      Vq dsv;
#pragma unroll
      for (int p = 0; p <= P; p++) dsv.x[p] = (p == 0 || bit(c.addm(), p)) ? st.x[p] - old.x[p] : 0.0;
      dsv = c.sum(dsv, decay_total<C>(c, S, WSI_SURF_N_DECAY));
      c.stv_add(S(WSI_SURF_MB_DS(V)), ss, dsv);
    }
    c.stv_add(S(WSI_SURF_MB_IN(V)), ss, mb_in);
    c.stv_add(S(WSI_SURF_MB_OUT(V)), ss, mb_out);
  }
  c.stv(stp, ss, st);
  if (growing) {
    *S(WSI_SURF_DAYS_AFTER_SOW) = days_after_sow;  // crop calendar state (calc_crop_cover)
    *S(WSI_SURF_CROP_COVER) = crop_cover;
    *S(WSI_SURF_GROUND_COVER) = ground_cover;
    *S(WSI_SURF_CROP_FACTOR) = crop_factor_out;
    *S(WSI_SURF_ET0_COEFFICIENT) = et0_coef;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      *S(WSI_SURF_POOL_FAST + k) = pl.fast[k];
      *S(WSI_SURF_POOL_HUMUS + k) = pl.humus[k];
      *S(WSI_SURF_POOL_DIS_INORG + k) = pl.dis_inorg[k];
      *S(WSI_SURF_POOL_DIS_ORG + k) = pl.dis_org[k];
      *S(WSI_SURF_POOL_ADS_INORG + k) = pl.ads_inorg[k];
    }
  }
  // per-step diagnostics recorded by Model.run (model.py:1081-1118); a GardenSurface's evaporation and
  // precipitation are also what next step's irrigation demand is worked out from (land.py:2402-2423)
  if (c.D().n_probe > 0 || (int)*Lo(WSI_LAND_GARDEN_SURF) == u || (c.D().features & 2)) {
    *S(WSI_SURF_OUT_PERCOLATION) = pc.x[0];
    *S(WSI_SURF_OUT_SUBSURFACE) = sf.x[0];
    *S(WSI_SURF_OUT_SURFACE) = type == WSI_ST_IMPERVIOUSSURFACE ? 0.0 : ie.x[0];
    *S(WSI_SURF_OUT_EVAPORATION) = out_evap;
    *S(WSI_SURF_OUT_PRECIPITATION) = out_precip;
    *S(WSI_SURF_OUT_TANK_RECHARGE) = out_recharge;
  }
}

// ResidenceTank.pull_outflow tanks.py:387-400 on a register-resident tank
template <class C>
WSI_HD typename C::Vq residence_pull(const C& c, typename C::Vq& s, double residence_time) {
  const double outflow = vol_vchange(s.x[0], div_z(s.x[0], residence_time));  // residence_time > 0
  return vq_pull(c, s, outflow);
}

// Mass-balance terms of a Land node (core.py:370-429 with land.py:113,158-162), assembled by
// the mass-balance pass from what is in memory after Land.run: the surfaces' shares (scratch planes), the out-arcs,
// and the three ResidenceTanks whose storage_ land_phase parked in the node's mb_terms row.
template <class C>
WSI_HD void land_mass_balance_terms(const C& c, int node, typename C::Vq& in_, typename C::Vq& ds, typename C::Vq& out_) {
  using Vq = typename C::Vq;
  constexpr int P = C::P;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int L = d.node_gidx[node];
  const size_t sl = c.s_land(), ss = c.s_surf();
  in_ = total_in(c, node);
  out_ = total_out(c, node);
  ds = vq_zero<P>();
  for (int i = d.land_surf_ptr[L]; i < d.land_surf_ptr[L + 1]; i++) {
    const Obj S = c.surf_obj(d.land_surf[i]);
    mb_acc(c, in_, c.ldv_add(S(WSI_SURF_MB_IN(V)), ss));
    mb_acc(c, out_, c.ldv_add(S(WSI_SURF_MB_OUT(V)), ss));
  }
  for (int i = d.land_surf_ptr[L]; i < d.land_surf_ptr[L + 1]; i++) {
    const Obj S = c.surf_obj(d.land_surf[i]);
    if ((int)*S(WSI_SURF_TYPE) == WSI_ST_IMPERVIOUSSURFACE) {
      // its tank changed again after surface_phase: ds from the parked storage_
      const Vq st = c.ldv_add(S(WSI_SURF_STORAGE(V)), ss);
      const Vq old = c.ldv_add(S(WSI_SURF_MB_DS(V)), ss);
      Vq dsv;
#pragma unroll
      for (int p = 0; p <= P; p++) dsv.x[p] = (p == 0 || bit(c.addm(), p)) ? st.x[p] - old.x[p] : 0.0;
      dsv = c.sum(dsv, decay_total<C>(c, S, WSI_SURF_N_DECAY));
      ds_add(c, ds, dsv);
    } else {
      ds_add(c, ds, c.ldv_add(S(WSI_SURF_MB_DS(V)), ss));
    }
  }
  // surface_runoff.ds, subsurface_runoff.ds, percolation.ds (land.py:160-162)
  const double* parked = d.mb_terms + (size_t)d.mb_term_slot[node] * 3 * V * c.R + c.r;
  const int tank_plane[3] = {WSI_LAND_SURFACE_RUNOFF(V), WSI_LAND_SUBSURFACE_RUNOFF(V), WSI_LAND_PERCOLATION(V)};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const Vq fin = c.ldv_add(c.land(tank_plane[k], L), sl);
    const Vq ini = c.ldv_add(parked + (size_t)k * V * c.R, c.R);
#pragma unroll
    for (int p = 0; p <= P; p++)
      if (p == 0 || bit(c.addm(), p)) ds.x[p] += fin.x[p] - ini.x[p];
  }
}

// Land.run land.py:200-244 after its surfaces have run (surface_phase), as sub-steps of the
// schedule entry (run_op, wsi_nodes.cuh): one per surface -- an ImperviousSurface's ponded runoff
// goes to the Sewers (push_to_sewers land.py:699-721, second half) -- then the percolation tank,
// then the two runoff tanks.  Every sub-step prepares at most one push (land_pre) and consumes its
// reply (land_post).  The surfaces' route() pushes (land.py:1046-1058) are applied tank by tank --
// each tank still receives them in surface order, so every sum is the reference's -- which lets
// each ResidenceTank be loaded and stored once.
struct OpAux {
  double x0, x1, x2;
  int i;
};

template <class C>
WSI_HD int land_substeps(const C& c, int node) {
  const WsiModelDesc& d = c.D();
  const int L = d.node_gidx[node];
  return d.land_surf_ptr[L + 1] - d.land_surf_ptr[L] + 2;
}

template <class C>
WSI_HD bool land_pre(C& c, int node, int sub, typename C::Vq& v, int& filter, OpAux& aux) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int L = d.node_gidx[node];
  const size_t sl = c.s_land();
  const size_t ss = c.s_surf();
  const bool want_mb = (c.run_flags & WSI_RUN_MASS_BALANCE) != 0;
  const int s0 = d.land_surf_ptr[L], s1 = d.land_surf_ptr[L + 1];
  double* parked = d.mb_terms + (size_t)d.mb_term_slot[node] * 3 * V * c.R + c.r;
  if (sub < s1 - s0) {  // impervious surfaces -> Sewers (independent of the Land's own tanks)
    const int u = d.land_surf[s0 + sub];
    if ((int)*c.surf(WSI_SURF_TYPE, u) != WSI_ST_IMPERVIOUSSURFACE) return false;
    v = c.ldv(c.surf(WSI_SURF_FLOW_A(V), u), ss);
    filter = WSI_FILTER_SEWER;
    aux.i = u;
    return true;
  }
  if (sub == s1 - s0) {  // percolation tank -> Groundwater (land.py:207-215)
    double* tp = c.land(WSI_LAND_PERCOLATION(V), L);
    Vq T = c.ldv(tp, sl);
    if (want_mb) c.stv_add(parked + (size_t)2 * V * c.R, c.R, T);  // storage_ for the mass balance
    for (int i = s0; i < s1; i++) {
      const Obj S = c.surf_obj(d.land_surf[i]);
      if ((int)*S(WSI_SURF_TYPE) >= WSI_ST_PERVIOUSSURFACE) T = c.sum(T, c.ldv(S(WSI_SURF_FLOW_C(V)), ss));
    }
    v = residence_pull(c, T, *c.land(WSI_LAND_RES_PERCOLATION, L));
    c.stv(tp, sl, T);
    filter = WSI_FILTER_GROUNDWATER;
    return true;
  }
  // surface + subsurface runoff -> River / junction Node (land.py:217-244)
  double* t_surf = c.land(WSI_LAND_SURFACE_RUNOFF(V), L);
  double* t_sub = c.land(WSI_LAND_SUBSURFACE_RUNOFF(V), L);
  {
    Vq T = c.ldv(t_surf, sl);
    if (want_mb) c.stv_add(parked, c.R, T);
    for (int i = s0; i < s1; i++) {
      const Obj S = c.surf_obj(d.land_surf[i]);
      if ((int)*S(WSI_SURF_TYPE) >= WSI_ST_PERVIOUSSURFACE) T = c.sum(T, c.ldv(S(WSI_SURF_FLOW_A(V)), ss));
    }
    v = residence_pull(c, T, *c.land(WSI_LAND_RES_SURFACE, L));
    aux.x0 = v.x[0];
    c.stv(t_surf, sl, T);
  }
  {
    Vq T = c.ldv(t_sub, sl);
    if (want_mb) c.stv_add(parked + (size_t)V * c.R, c.R, T);
    for (int i = s0; i < s1; i++) {
      const Obj S = c.surf_obj(d.land_surf[i]);
      if ((int)*S(WSI_SURF_TYPE) >= WSI_ST_PERVIOUSSURFACE) T = c.sum(T, c.ldv(S(WSI_SURF_FLOW_B(V)), ss));
    }
    const Vq subsurface_runoff = residence_pull(c, T, *c.land(WSI_LAND_RES_SUBSURFACE, L));
    aux.x1 = subsurface_runoff.x[0];
    c.stv(t_sub, sl, T);
    v = c.sum(v, subsurface_runoff);
  }
  aux.x2 = v.x[0];
  filter = WSI_FILTER_RIVER_NODE;
  return v.x[0] > 0;
}

template <class C>
WSI_HD void land_post(C& c, int node, int sub, const typename C::Vq& reply, const OpAux& aux) {
  using Vq = typename C::Vq;
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int L = d.node_gidx[node];
  const size_t sl = c.s_land();
  const int ns = d.land_surf_ptr[L + 1] - d.land_surf_ptr[L];
  if (sub < ns) {  // whatever the Sewers do not take is forced back into the surface tank
    tank_push_force(c, c.surf(WSI_SURF_STORAGE(V), aux.i), c.s_surf(), reply);
  } else if (sub == ns) {
    if (reply.x[0] > c.FA) tank_push_force(c, c.land(WSI_LAND_PERCOLATION(V), L), sl, reply);
  } else if (reply.x[0] > 0) {
    const double v_total = aux.x2;
    const Vq rs = c.vchange(reply, reply.x[0] * aux.x0 / v_total);
    const Vq rss = c.vchange(reply, reply.x[0] * aux.x1 / v_total);
    if (rs.x[0] > 0) tank_push_force(c, c.land(WSI_LAND_SURFACE_RUNOFF(V), L), sl, rs);
    if (rss.x[0] > 0) tank_push_force(c, c.land(WSI_LAND_SUBSURFACE_RUNOFF(V), L), sl, rss);
  }
}

// Land.end_timestep land.py:266-275: surfaces are DecayTanks (tanks.py:442-447); the
// residence tanks only refresh storage_, which is implicit here.
template <class C>
WSI_HD void land_end_timestep(C& c, int node) {
  constexpr int V = C::V;
  const WsiModelDesc& d = c.D();
  const int L = d.node_gidx[node];
  const double temperature = c.forcing((int)*c.land(WSI_LAND_FORCING_ID, L), 2);
  for (int i = d.land_surf_ptr[L]; i < d.land_surf_ptr[L + 1]; i++) {
    const int u = d.land_surf[i];
    const Obj S = c.surf_obj(u);
    if ((int)*S(WSI_SURF_N_DECAY) > 0) {
      double* stp = S(WSI_SURF_STORAGE(V));
      typename C::Vq st = c.ldv(stp, c.s_surf());
      decay_end_timestep(c, st, S, WSI_SURF_N_DECAY, temperature);
      c.stv(stp, c.s_surf(), st);
    }
  }
}

}  // namespace wsi
