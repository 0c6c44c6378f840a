/*
 * oracle/oracle_sunsky.cpp — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * The reference's procedural sun & sky environment (shaders/sun_and_sky.glsl; selected by SunAndSky.in_use == 1 in EnvRadiance /
 * EnvEval / EnvSample, pathtrace.glsl:40-72, env_sampling.glsl:111-125), restated for the CPU: one function per GLSL function (each
 * cites its lines), the Perez term that sky_color_xyz and sky_luminance share factored into one helper.
 *
 * Numerics: fp32, one rounding per written operation in the GLSL's evaluation order; exp / pow / acos / sin / cos from
 * include/eid_detmath.h, tan(x) = sin(x) / cos(x), smoothstep = t*t*(3-2t) on the clamped ratio.  Compiled with -ffp-contract=off.
 * Parity: PINNED — bit-identical to sun_and_sky.glsl itself compiled as C++ (oracle/ref_shim, tests/test_oracle_kat.py), given the
 * contract's built-ins.
 */
#include "oracle.h"

namespace orc {

static inline float ssLuminance(vec3 c) { return (0.2126f * c.x + 0.7152f * c.y) + 0.0722f * c.z; }   // :31-34
static inline vec3 ssV(const eid_vec3& v) { return vec3(v.x, v.y, v.z); }

static const float SS_PI = 3.1415926535f;   // sun_and_sky.glsl:26 (its own, shorter M_PI)

static inline float ssSmoothstep(float a, float b, float x) {
  float t = (x - a) / (b - a);
  t = gmin(gmax(t, 0.0f), 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
static inline vec3 ssExp3(vec3 v) { return vec3(eid_expf(v.x), eid_expf(v.y), eid_expf(v.z)); }
static inline vec3 ssPow3(vec3 v, float e) { return vec3(eid_powf(v.x, e), eid_powf(v.y, e), eid_powf(v.z, e)); }

// xyz2dir :37-71
static inline vec3 ssXyz2dir(vec3 m, float x, float y, float z) {
  vec3 u;
  if (fabsf(m.x) < fabsf(m.y)) u = vec3(0.0f, -m.z, m.y);
  else u = vec3(m.z, 0.0f, -m.x);
  // (the "degenerate transform" branch at :55-65 recomputes the same u from the same vector)
  u = normalize(u);
  const vec3 v = cross(m, u);
  return (x * u + y * v) + z * m;
}

// mi_lib_square_to_disk :74-115 -> (r, phi)
static inline void ssSquareToDisk(float inX, float inY, float& r, float& phi) {
  const float lx = 2.0f * inX - 1.0f, ly = 2.0f * inY - 1.0f;
  if (lx == 0.0f && ly == 0.0f) { phi = 0.0f; r = 0.0f; return; }
  if (lx > -ly) {
    if (lx > ly) { r = lx; phi = (SS_PI / 4.0f) * (1.0f + ly / lx); }
    else { r = ly; phi = (SS_PI / 4.0f) * (3.0f - lx / ly); }
  } else {
    if (lx < ly) { r = -lx; phi = (SS_PI / 4.0f) * (5.0f + ly / lx); }
    else { r = -ly; phi = (SS_PI / 4.0f) * (7.0f - lx / ly); }
  }
}

// mi_reflection_dir_diffuse_x :118-138
static inline vec3 ssDiffuseDir(vec3 normal, float sx, float sy) {
  float r, phi;
  ssSquareToDisk(sx, sy, r, phi);
  const float x = r * eid_cosf(phi), y = r * eid_sinf(phi);
  const float z2 = (1.0f - x * x) - y * y;
  const float z = z2 > 0.0f ? sqrtf(z2) : 0.0f;
  return ssXyz2dir(normal, x, y, z);
}

// calc_sun_color :141-164
static inline vec3 ssSunColor(vec3 sunDir, float turbidity) {
  vec3 sunColor = vec3(0.0f);
  const vec3 ko = vec3(12.0f, 8.5f, 0.9f);
  const vec3 wavelength = vec3(0.610f, 0.550f, 0.470f);
  const vec3 solRad = vec3(1.0f * 127500.0f / 0.9878f, 0.992f * 127500.0f / 0.9878f, 0.911f * 127500.0f / 0.9878f);
  if (sunDir.z > 0.0f) {
    const float m = 1.0f / (sunDir.z + 0.15f * eid_powf(93.885f - eid_acosf(sunDir.z) * 180.0f / SS_PI, -1.253f));
    const float beta = 0.04608f * turbidity - 0.04586f;
    const float alpha = 1.3f;
    const vec3 ta = ssExp3((-m * beta) * ssPow3(wavelength, -alpha));   // aerosol attenuation
    const float l = 0.0035f;
    const vec3 to = ssExp3(((-m) * ko) * l);                            // ozone absorption
    const vec3 tr = ssExp3((-m * 0.008735f) * ssPow3(wavelength, -4.08f));   // Rayleigh scattering
    sunColor = ((tr * ta) * to) * solRad;
  }
  return sunColor;
}

// the Perez term shared by sky_color_xyz and sky_luminance
static inline float ssPerez(float A, float B, float C, float D, float E, float cosTheta, float gamma, float cosGamma, float thetaSun, float cosThetaSun) {
  return ((1.0f + A * eid_expf(B / cosTheta)) * ((1.0f + C * eid_expf(D * gamma)) + (E * cosGamma) * cosGamma)) /
         ((1.0f + A * eid_expf(B / 1.0f)) * ((1.0f + C * eid_expf(D * thetaSun)) + (E * cosThetaSun) * cosThetaSun));
}

// sky_color_xyz :167-221
static inline vec3 ssSkyColorXyz(vec3 dir, vec3 sunPos, float T, float lum) {
  float cosGamma = dot(sunPos, dir);
  if (cosGamma > 1.0f) cosGamma = 2.0f - cosGamma;
  const float gamma = eid_acosf(cosGamma);
  const float cosTheta = dir.z, cosThetaSun = sunPos.z;
  const float thetaSun = eid_acosf(cosThetaSun);
  const float t2 = T * T, ts2 = thetaSun * thetaSun, ts3 = ts2 * thetaSun;
  const float zenithX = (((0.001650f * ts3 - 0.003742f * ts2) + 0.002088f * thetaSun) + 0.0f) * t2 +
                        (((-0.029028f * ts3 + 0.063773f * ts2) - 0.032020f * thetaSun) + 0.003948f) * T +
                        (((0.116936f * ts3 - 0.211960f * ts2) + 0.060523f * thetaSun) + 0.258852f);
  const float zenithY = (((0.002759f * ts3 - 0.006105f * ts2) + 0.003162f * thetaSun) + 0.0f) * t2 +
                        (((-0.042149f * ts3 + 0.089701f * ts2) - 0.041536f * thetaSun) + 0.005158f) * T +
                        (((0.153467f * ts3 - 0.267568f * ts2) + 0.066698f * thetaSun) + 0.266881f);
  float A = -0.019257f * T - (0.29f - eid_powf(cosThetaSun, 0.5f) * 0.09f);
  float B = -0.066513f * T + 0.000818f, C = -0.000417f * T + 0.212479f, D = -0.064097f * T - 0.898875f, E = -0.003251f * T + 0.045178f;
  float x = ssPerez(A, B, C, D, E, cosTheta, gamma, cosGamma, thetaSun, cosThetaSun);
  A = -0.016698f * T - 0.260787f; B = -0.094958f * T + 0.009213f; C = -0.007928f * T + 0.210230f; D = -0.044050f * T - 1.653694f;
  E = -0.010922f * T + 0.052919f;
  float y = ssPerez(A, B, C, D, E, cosTheta, gamma, cosGamma, thetaSun, cosThetaSun);
  const float sat = 1.0f;
  x = zenithX * (x * sat + (1.0f - sat));
  y = zenithY * (y * sat + (1.0f - sat));
  vec3 xyz;
  xyz.y = lum;
  xyz.x = (x / y) * xyz.y;
  xyz.z = (((1.0f - x) - y) / y) * xyz.y;
  return xyz;
}

// sky_luminance :224-250
static inline float ssSkyLuminance(vec3 dir, vec3 sunPos, float T) {
  float cosGamma = dot(sunPos, dir);
  if (cosGamma < 0.0f) cosGamma = 0.0f;
  if (cosGamma > 1.0f) cosGamma = 2.0f - cosGamma;
  const float gamma = eid_acosf(cosGamma);
  const float cosTheta = dir.z, cosThetaSun = sunPos.z;
  const float thetaSun = eid_acosf(cosThetaSun);
  const float A = 0.178721f * T - 1.463037f, B = -0.355402f * T + 0.427494f, C = -0.022669f * T + 5.325056f;
  const float D = 0.120647f * T - 2.577052f, E = -0.066967f * T + 0.370275f;
  return ssPerez(A, B, C, D, E, cosTheta, gamma, cosGamma, thetaSun, cosThetaSun);
}

// calc_env_color :253-266
static inline vec3 ssEnvColor(vec3 sunDir, vec3 dir, float T) {
  const float thetaSun = eid_acosf(sunDir.z);
  const float chi = (4.0f / 9.0f - T / 120.0f) * (SS_PI - 2.0f * thetaSun);
  float s, c;
  eid_sincosf(chi, &s, &c);
  float lum = 1000.0f * (((4.0453f * T - 4.9710f) * (s / c) - 0.2155f * T) + 2.4192f);
  lum = lum * ssSkyLuminance(dir, sunDir, T);
  const vec3 X = ssSkyColorXyz(dir, sunDir, T, lum);
  vec3 env = vec3((3.241f * X.x - 1.537f * X.y) - 0.499f * X.z, (-0.969f * X.x + 1.876f * X.y) + 0.042f * X.z, (0.056f * X.x - 0.204f * X.y) + 1.057f * X.z);
  return env * SS_PI;
}

// calc_irrad :269-289 (5 x 5 stratified directions of the upper hemisphere; float loop counters as in the GLSL)
static inline vec3 ssIrrad(vec3 sunDir, float haze) {
  vec3 acc = vec3(0.0f);
  const vec3 n = vec3(0.0f, 0.0f, 1.0f);
  for (float u = 1.0f / 10.0f; u < 1.0f; u += 1.0f / 5.0f)
    for (float v = 1.0f / 10.0f; v < 1.0f; v += 1.0f / 5.0f)
      acc = acc + ssEnvColor(sunDir, ssDiffuseDir(n, u, v), haze);
  return acc / 25.0f;
}

// tweak_saturation :292-308
static inline float ssTweakSaturation(float sat, float haze) {
  const float lowsat = eid_powf(sat, 3.0f);
  if (sat <= 1.0f) {
    float h = haze;
    h = h - 2.0f;
    h = h / 15.0f;
    if (h < 0.0f) h = 0.0f;
    if (h > 1.0f) h = 1.0f;
    h = eid_powf(h, 3.0f);
    return sat * (1.0f - h) + lowsat * h;
  }
  return 1.0f;
}

// arch_vectortweak :311-324
static inline vec3 ssVectorTweak(vec3 dir, int yIsUp, float horizHeight) {
  vec3 o = dir;
  if (yIsUp == 1) o = vec3(dir.x, dir.z, dir.y);
  if (horizHeight != 0.0f) { o.z = o.z - horizHeight; o = normalize(o); }
  return o;
}

// arch_colortweak :327-356 (the clamp of negative components at :342-352 only touches a dead copy of `tint`)
static inline vec3 ssColorTweak(vec3 tint, float saturation, float redness) {
  const float intensity = ssLuminance(tint);
  vec3 o;
  if (saturation <= 0.0f) o = vec3(intensity);
  else o = tint * saturation + vec3(intensity * (1.0f - saturation));
  return o * vec3(1.0f + redness, 1.0f, 1.0f - redness);
}

// calc_physical_scale :359-438 -> (sun disk scale, sun glow scale)
static inline void ssPhysicalScale(float diskScale, float glowIntensity, float diskIntensity, float& sundiskScale, float& sunglowScale) {
  const float sunAngularRadius = 0.00465f;
  const float diskRadius = sunAngularRadius * diskScale;
  const float glowRadius = diskRadius * 10.0f;
  const float glowIntegral = glowIntensity * (((4.0f * SS_PI) - (24.0f * SS_PI) / (glowRadius * glowRadius)) +
                                              ((24.0f * SS_PI) * eid_sinf(glowRadius)) / ((glowRadius * glowRadius) * glowRadius));
  float target = diskIntensity * SS_PI;
  sunglowScale = 1.0f;
  const float maxGlow = 0.5f * target;
  if (glowIntegral > maxGlow) { sunglowScale = sunglowScale * (maxGlow / glowIntegral); target = target - maxGlow; }
  else target = target - glowIntegral;
  const float area = (2.0f * SS_PI) * (1.0f - eid_cosf(diskRadius));
  const float targetIntensity = target / area;
  const float actualIntegral = 1.0f * area;
  const float actualIntensity = ((diskIntensity * 100.0f) * actualIntegral) / area;
  sundiskScale = (targetIntensity == 0.0f) ? 0.0f : targetIntensity / actualIntensity;
}

// night_brightness_adjustment :441-450
static inline float ssNightBrightness(vec3 sunDir) {
  const float lmt = 0.30901699437494742410229341718282f;
  if (sunDir.z <= -lmt) return 0.0f;
  float f = (sunDir.z + lmt) / lmt;
  f = f * f;
  f = f * f;
  return f;
}

// sun_and_sky :453-601
vec3 sun_and_sky(const SunAndSky& ss, vec3 inDirection) {
  float factor = 1.0f, nightFactor = 1.0f;
  vec3 rgbScale = ssV(ss.rgb_unit_conversion);
  const float horizHeight = ss.horizon_height / 10.0f;
  vec3 dir = ssVectorTweak(inDirection, ss.y_is_up, horizHeight);
  float localHaze = 2.0f + ss.haze;
  if (localHaze < 2.0f) localHaze = 2.0f;
  const float localSaturation = ssTweakSaturation(ss.saturation, localHaze);
  if (ssLuminance(rgbScale) < 0.0f) rgbScale = vec3(1.0f / 80000.0f);
  rgbScale = rgbScale * ss.multiplier;
  if (ss.multiplier <= 0.0f) return vec3(0.0f);

  const float downness = dir.z;
  const vec3 realDir = dir;
  if (dir.z < 0.001f) { dir.z = 0.001f; dir = normalize(dir); }   // only calc for above-the-horizon

  vec3 sunDir = normalize(ssV(ss.sun_direction));
  sunDir = ssVectorTweak(sunDir, ss.y_is_up, horizHeight);
  const vec3 realSunDir = sunDir;
  if (sunDir.z < 0.001f) {
    if (sunDir.z < 0.0f) factor = ssNightBrightness(sunDir);
    sunDir.z = 0.001f;
    sunDir = normalize(sunDir);
  }

  vec3 tint;
  if (factor > 0.0f) {
    tint = ssEnvColor(sunDir, dir, localHaze);
    if (factor < 1.0f) tint = tint * factor;
  } else tint = vec3(0.0f);
  const vec3 sunColor = ssSunColor(sunDir, downness > 0.0f ? localHaze : 2.0f);
  if (ss.sun_disk_intensity > 0.0f && ss.sun_disk_scale > 0.0f) {
    const float sunAngle = eid_acosf(dot(realDir, realSunDir));
    const float sunRadius = (0.00465f * ss.sun_disk_scale) * 10.0f;
    if (sunAngle < sunRadius) {
      float sundiskScale = 1.0f, sunglowScale = 1.0f;
      if (ss.physically_scaled_sun == 1) ssPhysicalScale(ss.sun_disk_scale, ss.sun_glow_intensity, ss.sun_disk_intensity, sundiskScale, sunglowScale);
      float sunFactor = (1.0f - sunAngle / sunRadius) * 10.0f;
      sunFactor = ((eid_powf(sunFactor / 10.0f, 3.0f) * 2.0f) * ss.sun_glow_intensity) * sunglowScale +
                  ((ssSmoothstep(8.5f, 9.5f + (localHaze / 50.0f), sunFactor) * 100.0f) * ss.sun_disk_intensity) * sundiskScale;
      tint = tint + sunColor * sunFactor;
    }
  }
  vec3 outColor = tint * rgbScale;
  if (downness <= 0.0f) {
    vec3 downColor = ssV(ss.ground_color);
    const vec3 irrad = ssIrrad(sunDir, 2.0f);
    downColor = downColor * ((irrad + sunColor * sunDir.z) * rgbScale);
    if (factor < 1.0f) downColor = downColor * factor;
    const float horBlur = ss.horizon_blur / 10.0f;
    if (horBlur > 0.0f) {
      float dness = -downness;
      dness = dness / horBlur;
      if (dness > 1.0f) dness = 1.0f;
      dness = ssSmoothstep(0.0f, 1.0f, dness);
      outColor = outColor * (1.0f - dness) + downColor * dness;
      nightFactor = 1.0f - dness;
    } else { outColor = downColor; nightFactor = 0.0f; }
  }
  vec3 result = ssColorTweak(outColor, localSaturation, ss.redblueshift);
  if (nightFactor > 0.0f) {
    const vec3 night = ssV(ss.night_color) * nightFactor;
    if (result.x < night.x) result.x = night.x;
    if (result.y < night.y) result.y = night.y;
    if (result.z < night.z) result.z = night.z;
  }
  return result * SS_PI;
}

}  // namespace orc
