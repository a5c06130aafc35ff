// tuner.cpp -- the negative-feedback compaction policy (host side, FP64).
//
// CompactTuner (negative_feedback.hpp:165-260): one UCB1-tuned bandit per join whose
// arms are compaction thresholds.  MultiArmedBandit (negative_feedback.hpp:20-163):
// forced round-robin warm-up (kArms * 4 pulls), EMA reward / reward^2 with a window of
// 15, UCB-tuned exploration bonus, and a restart when an arm's estimate drifts by more
// than 2x between two heart beats (every 256 selections).
// The arithmetic is kept operation-for-operation identical (same expression order, so
// the doubles agree bit for bit with the reference given the same reward sequence);
// the policy is pure host code -- rewards come from device-side timing of the fused
// chain kernel (cc_chain_result.cycles), thresholds go back as kernel arguments.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <sys/stat.h>

#include "cc_api.h"

namespace ccb {
void set_error(const char *fmt, ...);
}

namespace {

class Bandit {
 public:
  explicit Bandit(size_t n_arms)
      : arms_(n_arms), est_(n_arms, 0.0), est_sq_(n_arms, 0.0), n_select_(n_arms, 0), stage_n_update_(n_arms, 0) {}

  size_t Select() {  // negative_feedback.hpp:34-61
    if (n_start_sampling_ < arms_ * kStartSampling) {
      size_t arm = n_start_sampling_ % arms_;
      ++n_start_sampling_;
      ++select_times_;
      ++n_select_[arm];
      return arm;
    }
    double best = -1;
    size_t best_arm = 0;
    for (size_t i = 0; i < arms_; ++i) {
      double value = est_[i] + UcbTuned(i);
      if (value > best) {
        best = value;
        best_arm = i;
      }
    }
    ++select_times_;
    ++n_select_[best_arm];
    return best_arm;
  }

  void Update(size_t arm, double reward) {  // negative_feedback.hpp:64-91
    if (select_times_ % kHeart == 0 && n_start_sampling_ >= arms_ * kStartSampling) {
      history_rewards_.push_back(est_);
      history_selects_.push_back(n_select_);
      if (r_means_.empty()) r_means_ = est_;
      bool drift = est_[arm] > r_means_[arm] * 2 || est_[arm] < r_means_[arm] / 2;
      r_means_ = est_;
      if (drift) {
        n_start_sampling_ = 0;
        std::fill(est_.begin(), est_.end(), 0.0);
        std::fill(est_sq_.begin(), est_sq_.end(), 0.0);
        stage_update_times_ = 0;
        std::fill(stage_n_update_.begin(), stage_n_update_.end(), size_t(0));
      }
    }
    size_t update_factor = std::min(stage_n_update_[arm], size_t(15));
    double ratio = update_factor / (update_factor + 1.0);
    est_[arm] = est_[arm] * ratio + reward * (1 - ratio);
    est_sq_[arm] = est_sq_[arm] * ratio + reward * reward * (1 - ratio);
    ++stage_update_times_;
    ++stage_n_update_[arm];
  }

  bool LogCsv(const std::string &path) const {  // negative_feedback.hpp:99-120
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return false;
    for (size_t i = 0; i < history_rewards_.size(); ++i) {
      fprintf(f, "%zu, ", i * kHeart);
      for (double r : history_rewards_[i]) fprintf(f, "%g, ", r);
      for (size_t s : history_selects_[i]) fprintf(f, "%zu, ", s);
      fprintf(f, "\n");
    }
    fclose(f);
    return true;
  }

  const std::vector<double> &rewards() const { return est_; }
  const std::vector<size_t> &selects() const { return n_select_; }

 private:
  double UcbTuned(size_t arm) const {  // negative_feedback.hpp:123-127
    double ucb_var = est_sq_[arm] - est_[arm] * est_[arm] + sqrt(2 * log(stage_update_times_) / (stage_n_update_[arm] + kEpsilon));
    return sqrt(log(stage_update_times_) / (stage_n_update_[arm] + kEpsilon) * std::min(0.25, ucb_var));
  }

  static constexpr double kEpsilon = 0.1;
  static constexpr size_t kStartSampling = 4;
  static constexpr size_t kHeart = 256;
  size_t arms_;
  std::vector<double> est_, est_sq_;
  std::vector<size_t> n_select_, stage_n_update_;
  size_t select_times_ = 0, stage_update_times_ = 0, n_start_sampling_ = 0;
  std::vector<double> r_means_;
  std::vector<std::vector<double>> history_rewards_;
  std::vector<std::vector<size_t>> history_selects_;
};

struct Package {  // BanditPackage, negative_feedback.hpp:240-252
  std::unique_ptr<Bandit> bandit;
  std::vector<size_t> value;
  std::unordered_map<size_t, size_t> value_index;
  size_t address;
};

}  // namespace

struct cc_tuner {
  std::unordered_map<size_t, size_t> package_index;
  std::vector<Package> packages;
};

extern "C" {

int cc_tuner_create(cc_tuner **t) {
  if (!t) {
    ccb::set_error("tuner is NULL");
    return CC_ERR_INVALID;
  }
  *t = new cc_tuner();
  return CC_OK;
}

int cc_tuner_initialize(cc_tuner *t, size_t address, const size_t *arms, size_t n_arms) {
  static const size_t kDefaultArms[] = {0, 32, 64, 128, 256, 384, 512, 768, 1024};  // negative_feedback.hpp:172
  if (!t) {
    ccb::set_error("tuner is NULL");
    return CC_ERR_INVALID;
  }
  if (t->package_index.count(address)) {  // assert at :173
    ccb::set_error("address %zu already registered", address);
    return CC_ERR_STATE;
  }
  if (!arms) {
    arms = kDefaultArms;
    n_arms = sizeof(kDefaultArms) / sizeof(kDefaultArms[0]);
  }
  if (n_arms == 0) {
    ccb::set_error("a bandit needs at least one arm");
    return CC_ERR_INVALID;
  }
  Package p;
  p.bandit = std::make_unique<Bandit>(n_arms);
  p.value.assign(arms, arms + n_arms);
  for (size_t i = 0; i < n_arms; ++i) p.value_index[arms[i]] = i;
  p.address = address;
  t->package_index[address] = t->packages.size();
  t->packages.push_back(std::move(p));
  return CC_OK;
}

int cc_tuner_select_arm(cc_tuner *t, size_t id, size_t *arm_value) {
  if (!t || !arm_value || id >= t->packages.size()) {
    ccb::set_error("bad tuner / bandit id %zu", id);
    return CC_ERR_INVALID;
  }
  Package &p = t->packages[id];
  *arm_value = p.value[p.bandit->Select()];
  return CC_OK;
}

int cc_tuner_update_arm(cc_tuner *t, size_t id, size_t arm_value, double reward) {
  if (!t || id >= t->packages.size()) {
    ccb::set_error("bad tuner / bandit id %zu", id);
    return CC_ERR_INVALID;
  }
  Package &p = t->packages[id];
  auto it = p.value_index.find(arm_value);
  if (it == p.value_index.end()) return CC_OK;  // unknown value is ignored (:193)
  p.bandit->Update(it->second, reward);
  return CC_OK;
}

int64_t cc_tuner_get_id(const cc_tuner *t, size_t address) {
  if (!t) return -1;
  auto it = t->package_index.find(address);
  return it == t->package_index.end() ? -1 : (int64_t) it->second;
}

size_t cc_tuner_bandit_size(const cc_tuner *t) { return t ? t->packages.size() : 0; }

int cc_tuner_state(const cc_tuner *t, size_t id, double *est_rewards, uint64_t *n_select, size_t n_arms) {
  if (!t || id >= t->packages.size()) {
    ccb::set_error("bad tuner / bandit id %zu", id);
    return CC_ERR_INVALID;
  }
  const Package &p = t->packages[id];
  if (n_arms != p.value.size()) {
    ccb::set_error("bandit %zu has %zu arms, caller passed %zu", id, p.value.size(), n_arms);
    return CC_ERR_INVALID;
  }
  for (size_t i = 0; i < n_arms; ++i) {
    if (est_rewards) est_rewards[i] = p.bandit->rewards()[i];
    if (n_select) n_select[i] = p.bandit->selects()[i];
  }
  return CC_OK;
}

int cc_tuner_reset(cc_tuner *t, int enable_log, const char *log_dir) {
  if (!t) {
    ccb::set_error("tuner is NULL");
    return CC_ERR_INVALID;
  }
  // the reference only clears its bandits when logging is enabled (:198-219)
  if (t->packages.empty() || !enable_log) return CC_OK;
  std::string dir = log_dir && *log_dir ? log_dir : "./bandit_log";
  mkdir(dir.c_str(), 0755);
  for (auto &kv : t->package_index) {
    const Package &p = t->packages[kv.second];
    std::string name = dir + "/0x" + std::to_string(kv.first) + "_Id-" + std::to_string(kv.second) + ".log";
    if (!p.bandit->LogCsv(name)) {
      ccb::set_error("Unable to open file %s", name.c_str());  // :103-105
      return CC_ERR_INVALID;
    }
    for (size_t i = 0; i < p.value.size(); ++i)
      fprintf(stderr, " [PARAMETERS] Estimated reward for arm %zu is %f - Sampling times is %zu\n", p.value[i],
              p.bandit->rewards()[i], p.bandit->selects()[i]);
  }
  t->package_index.clear();
  t->packages.clear();
  return CC_OK;
}

int cc_tuner_destroy(cc_tuner *t) {
  delete t;
  return CC_OK;
}

}  // extern "C"
