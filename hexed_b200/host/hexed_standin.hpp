/* hexed_standin.hpp -- header-compatible stand-ins for the slice of the Hexed API that the adapter (adapter.cpp) touches.
 *
 * The genuine headers (include/kernels.hpp, Kernel_mesh.hpp, Basis.hpp, Transport_model.hpp, Face_permutation.hpp of the Hexed
 * tree) need Eigen, which is not installed in this image, so the adapter cannot be compiled against them here. This file
 * declares the SAME names with the SAME member signatures (only what adapter.cpp uses), with `hexed::Plain_mat` taking the
 * place of Eigen::MatrixXd / VectorXd as a return type: `m(i, j)`, `v(i)`, `rows()`, `cols()` work on both, and that is all
 * the adapter calls. A Hexed build defines HEXED_B200_WITH_HEXED_HEADERS and includes the genuine <kernels.hpp> instead
 * (INTEGRATION.md); nothing in adapter.cpp changes.
 *
 * Interfaces mirrored (reference file:line):
 *   Sequence<T>            include/Sequence.hpp:13-19          Kernel_element       include/Kernel_element.hpp:15-45
 *   Connection_direction   include/Kernel_connection.hpp:7-37  Kernel_connection    include/Kernel_connection.hpp:39-55
 *   Refined_face           include/Refined_face.hpp:9-15       Stopwatch(_tree)     include/Stopwatch.hpp:13-45, Stopwatch_tree.hpp:15-32
 *   Basis                  include/Basis.hpp:16-66             Transport_model      include/Transport_model.hpp:14-48
 *   Kernel_mesh            include/Kernel_mesh.hpp:14-25       Kernel_options + free functions   include/kernels.hpp:11-42
 *   Face_permutation_dynamic  include/Face_permutation.hpp:13-19
 */
#ifndef HEXED_B200_STANDIN_HPP_
#define HEXED_B200_STANDIN_HPP_

#include <array>
#include <chrono>
#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace hexed
{

template <typename T> class Sequence
{
  public:
  virtual int size() = 0;
  virtual T operator[](int index) = 0;
};

class Kernel_element
{
  public:
  virtual double* state() = 0;
  virtual double* residual_cache() = 0;
  virtual double* time_step_scale() = 0;
  virtual double& vertex_time_step_scale(int i_vertex) = 0;
  virtual double nominal_size() = 0;
  virtual double* face(int i_face, bool is_ldg) = 0;
  virtual bool deformed() const = 0;
  virtual double* reference_level_normals() = 0;
  virtual double* jacobian_determinant() = 0;
  virtual double* kernel_face_normal(int i_face) = 0;
  virtual double& uncert() = 0;
};

class Connection_direction
{
  public:
  std::array<int, 2> i_dim;
  std::array<bool, 2> face_sign;
  int i_face(int i_side) {return 2*i_dim[i_side] + face_sign[i_side];}
  bool flip_normal(int i_side) {return face_sign[i_side] == i_side;}
  bool flip_tangential() {return (i_dim[0] != i_dim[1]) && (flip_normal(0) == flip_normal(1));}
  bool transpose() {return (i_dim[0] == 0 && i_dim[1] == 2) || (i_dim[0] == 2 && i_dim[1] == 0);}
};

class Connection
{
  public:
  virtual Connection_direction get_direction() = 0;
};

class Kernel_connection : virtual public Connection
{
  public:
  virtual double* state(int i_side, bool is_ldg) = 0;
  virtual double* normal() = 0;
};

class Refined_face
{
  public:
  double* coarse = nullptr;
  std::array<double*, 4> fine {};
  std::array<bool, 2> stretch;
};

class Stopwatch
{
  int n = 0;
  double t = 0.;
  bool r = false;
  std::chrono::steady_clock::time_point time_started;
  public:
  class Operator
  {
    Stopwatch& sw;
    public:
    Operator(Stopwatch& stopwatch) : sw{stopwatch} {sw.start();}
    ~Operator() {sw.pause();}
  };
  void start()
  {
    if (r) throw std::runtime_error("stopwatch already running");
    r = true; time_started = std::chrono::steady_clock::now();
  }
  void pause()
  {
    if (!r) throw std::runtime_error("stopwatch not running");
    t += std::chrono::duration<double>(std::chrono::steady_clock::now() - time_started).count();
    r = false; ++n;
  }
  void reset() {t = 0.; n = 0;}
  bool running() const {return r;}
  int n_calls() const {return n;}
  double time() const {return t;}
};

class Stopwatch_tree
{
  public:
  Stopwatch stopwatch;
  std::map<std::string, Stopwatch_tree> children;
  int work_units_completed = 0;
  std::string work_unit_name;
  Stopwatch_tree(std::string work_unit_name_arg, std::map<std::string, Stopwatch_tree> init_children = {})
  : children{init_children}, work_unit_name{work_unit_name_arg} {}
};

//! what the adapter needs of an Eigen dense object: sizes and element access
class Plain_mat
{
  int _rows = 0, _cols = 0;
  std::vector<double> _data; // column-major like Eigen's default
  public:
  Plain_mat() = default;
  Plain_mat(int n_rows, int n_cols = 1) : _rows{n_rows}, _cols{n_cols}, _data(size_t(n_rows)*n_cols, 0.) {}
  int rows() const {return _rows;}
  int cols() const {return _cols;}
  int size() const {return _rows*_cols;}
  double& operator()(int i, int j) {return _data[size_t(j)*_rows + i];}
  double operator()(int i, int j) const {return _data[size_t(j)*_rows + i];}
  double& operator()(int i) {return _data[i];}
  double operator()(int i) const {return _data[i];}
};

class Basis
{
  protected:
  virtual double min_eig_convection() const = 0;
  virtual double quadratic_safety() const = 0;
  public:
  const int row_size;
  Basis(int row_size_arg) : row_size{row_size_arg} {}
  virtual ~Basis() = default;
  virtual double node(int i) const = 0;
  virtual Plain_mat node_weights() const = 0;
  virtual Plain_mat diff_mat() const = 0;
  virtual Plain_mat boundary() const = 0;
  virtual Plain_mat orthogonal(int degree) const = 0;
  virtual Plain_mat filter() const = 0;
  virtual Plain_mat prolong(int i_half) const = 0;
  virtual Plain_mat restrict(int i_half) const = 0;
  double max_cfl() const {return -2*quadratic_safety()/min_eig_convection();}
  double step_ratio() const {return .5/quadratic_safety();}
  virtual double min_eig_diffusion() const = 0;
};

class Transport_model
{
  double const_val;
  double ref_val;
  double ref_temp;
  double sqrt_ref_temp;
  double temp_offset;
  Transport_model(double cv, double rv, double rt, double to, bool iv)
  : const_val{cv}, ref_val{rv}, ref_temp{rt}, sqrt_ref_temp{std::sqrt(rt)}, temp_offset{to}, is_viscous{iv} {}
  public:
  const bool is_viscous;
  double coefficient(double sqrt_temp) const
  {
    double r = sqrt_temp/sqrt_ref_temp;
    return const_val + ref_val*r*r*r*(ref_temp + temp_offset)/(sqrt_temp*sqrt_temp + temp_offset);
  }
  static inline Transport_model inviscid() {return {0., 0., 1., 1., 0};}
  static inline Transport_model constant(double value) {return {value, 0., 1., 1., 1};}
  static inline Transport_model sutherland(double reference_value, double reference_temperature, double temperature_offset)
  {
    return {0., reference_value, reference_temperature, temperature_offset, 1};
  }
};

struct Kernel_mesh
{
  int n_dim;
  int row_size;
  const Basis& basis;
  Sequence<Kernel_connection&>& car_cons;
  Sequence<Kernel_connection&>& def_cons;
  Sequence<Kernel_element&>& car_elems;
  Sequence<Kernel_element&>& def_elems;
  Sequence<Kernel_element&>& elems;
  Sequence<Refined_face&>& ref_faces;
};

struct Kernel_options
{
  Stopwatch_tree& sw_car;
  Stopwatch_tree& sw_def;
  Stopwatch_tree& sw_pr;
  double dt;
  int i_stage;
  bool compute_residual = false;
  bool use_filter = false;
};

class Face_permutation_dynamic
{
  public:
  virtual ~Face_permutation_dynamic() = default;
  virtual void match_faces() = 0;
  virtual void restore() = 0;
};

void compute_euler(Kernel_mesh, Kernel_options);
void compute_advection(Kernel_mesh, Kernel_options, double advect_length);
void compute_navier_stokes(Kernel_mesh, Kernel_options, std::function<void()> flux_bc, Transport_model visc, Transport_model therm_cond);
void compute_smooth_av(Kernel_mesh, Kernel_options, std::function<void()> flux_bc, double diff_time, double chebyshev_step);
void compute_fix_therm_admis(Kernel_mesh, Kernel_options, std::function<void()> flux_bc);
double max_dt_euler(Kernel_mesh, Kernel_options, double convective_safety, double diffusive_safety, bool local_time);
double max_dt_navier_stokes(Kernel_mesh, Kernel_options, double convective_safety, double diffusive_safety, bool local_time,
                            Transport_model visc, Transport_model therm_cond);
double max_dt_advection(Kernel_mesh, Kernel_options, double convective_safety, double diffusive_safety, bool local_time, double advect_length);
double max_dt_smooth_av(Kernel_mesh, Kernel_options, double convective_safety, double diffusive_safety, bool local_time);
double max_dt_fix_therm_admis(Kernel_mesh, Kernel_options, double convective_safety, double diffusive_safety, bool local_time);
void compute_prolong(Kernel_mesh, bool scale = false, bool offset = false);
void compute_restrict(Kernel_mesh, bool scale = true, bool offset = false);
void compute_prolong_advection(Kernel_mesh);
std::unique_ptr<Face_permutation_dynamic> face_permutation(int n_dim, int row_size, Connection_direction, double* data);
void compute_write_face(Kernel_mesh);
void compute_write_face_advection(Kernel_mesh);
void compute_write_face_smooth_av(Kernel_mesh);
void stabilizing_art_visc(Kernel_mesh, double char_speed);

}
#endif
