// ls2d_boss.h -- the slice of srrg2_core's config system (srrg_config + srrg_boss) that the hot path is
// reached through: typed PARAM properties, classes registered by name, a manager that instantiates a
// BOSS-text configuration ("ClassName" { "#id": n, "name": ..., params..., {"#pointer": id} }) and resolves
// the links.  Written from the reference's USE of that API -- R/ = /root/reference/srrg2_laser_slam_2d/src/
// srrg2_laser_slam_2d/:
//   PARAM(...)               R/registration/correspondence_finder_projective_2d.h:16-26
//   BOSS_REGISTER_CLASS      R/instances.cpp:27-36
//   ConfigurableManager      apps/slam_app.cpp:39-53 (read, getByName), :87-167 (create, setValue, write)
// so that the reference's two configuration files load unchanged.
#pragma once

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace srrg2_core {

  // ------------------------------------------------------------------ parsed value tree (BOSS text ~ JSON)
  struct BossValue {
    enum Kind { Null, Number, String, Array, Object } kind = Null;
    double number = 0;
    std::string text;
    std::vector<BossValue> items;                             // Array
    std::vector<std::pair<std::string, BossValue>> fields;    // Object, in file order
    const BossValue* find(const std::string& key) const {
      for (const auto& f : fields) {
        if (f.first == key) return &f.second;
      }
      return nullptr;
    }
    bool isPointer() const { return kind == Object && find("#pointer"); }
    int pointerId() const { return (int) find("#pointer")->number; }
  };

  class Configurable;
  using ConfigurablePtr = std::shared_ptr<Configurable>;

  // ------------------------------------------------------------------ properties
  class PropertyBase {
  public:
    PropertyBase(const std::string& name, const std::string& doc, Configurable* owner, bool* changed_flag);
    virtual ~PropertyBase() {}
    const std::string& name() const { return _name; }
    const std::string& doc() const { return _doc; }
    // set from a parsed value; pointers are resolved through `resolve` (id -> object, -1 -> null)
    virtual void fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>& resolve) = 0;
    virtual void toText(std::ostream& os, const std::function<int(const Configurable*)>& idOf) const = 0;

  protected:
    void touch() {
      if (_changed_flag) *_changed_flag = true;
    }
    std::string _name, _doc;
    bool* _changed_flag;
  };

  template <typename T>
  class Property_ : public PropertyBase {
  public:
    Property_(const std::string& name, const std::string& doc, Configurable* owner, const T& def,
              bool* changed_flag = nullptr) :
      PropertyBase(name, doc, owner, changed_flag), _value(def) {}
    const T& value() const { return _value; }
    void setValue(const T& v) {
      _value = v;
      touch();
    }
    void fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>&) override;
    void toText(std::ostream& os, const std::function<int(const Configurable*)>&) const override;

  protected:
    T _value;
  };
  using PropertyFloat       = Property_<float>;
  using PropertyInt         = Property_<int>;
  using PropertyUnsignedInt = Property_<unsigned>;
  using PropertyBool        = Property_<bool>;
  using PropertyString      = Property_<std::string>;

  template <typename T>
  class PropertyVector_ : public PropertyBase {
  public:
    PropertyVector_(const std::string& name, const std::string& doc, Configurable* owner,
                    bool* changed_flag = nullptr) :
      PropertyBase(name, doc, owner, changed_flag) {}
    const std::vector<T>& value() const { return _value; }
    size_t size() const { return _value.size(); }
    const T& value(size_t i) const { return _value.at(i); }
    void pushBack(const T& v) {
      _value.push_back(v);
      touch();
    }
    void setValue(const std::vector<T>& v) {
      _value = v;
      touch();
    }
    void fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>&) override {
      if (v.kind != BossValue::Array) throw std::runtime_error("property " + _name + ": expected an array");
      _value.clear();
      for (const auto& i : v.items) _value.push_back((T) i.number);
      touch();
    }
    void toText(std::ostream& os, const std::function<int(const Configurable*)>&) const override {
      os << "[ ";
      for (size_t i = 0; i < _value.size(); ++i) os << (i ? ", " : "") << _value[i];
      os << " ]";
    }

  protected:
    std::vector<T> _value;
  };

  // link to another configurable (shared, may be null): PropertyConfigurable_<T>
  template <typename T>
  class PropertyConfigurable_ : public PropertyBase {
  public:
    PropertyConfigurable_(const std::string& name, const std::string& doc, Configurable* owner,
                          std::shared_ptr<T> def = nullptr, bool* changed_flag = nullptr) :
      PropertyBase(name, doc, owner, changed_flag), _value(def) {}
    const std::shared_ptr<T>& value() const { return _value; }
    void setValue(const std::shared_ptr<T>& v) {
      _value = v;
      touch();
    }
    T* operator->() const { return _value.get(); }
    void fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>& resolve) override {
      if (!v.isPointer()) throw std::runtime_error("property " + _name + ": expected a #pointer");
      ConfigurablePtr p = resolve(v.pointerId());
      std::shared_ptr<T> typed = std::dynamic_pointer_cast<T>(p);
      if (p && !typed) throw std::runtime_error("property " + _name + ": linked object has the wrong class");
      _value = typed;
      touch();
    }
    void toText(std::ostream& os, const std::function<int(const Configurable*)>& idOf) const override {
      os << "{ \"#pointer\" : " << (_value ? idOf((const Configurable*) _value.get()) : -1) << " }";
    }

  protected:
    std::shared_ptr<T> _value;
  };

  template <typename T>
  class PropertyConfigurableVector_ : public PropertyBase {
  public:
    PropertyConfigurableVector_(const std::string& name, const std::string& doc, Configurable* owner,
                                bool* changed_flag = nullptr) :
      PropertyBase(name, doc, owner, changed_flag) {}
    size_t size() const { return _value.size(); }
    const std::shared_ptr<T>& value(size_t i) const { return _value.at(i); }
    const std::vector<std::shared_ptr<T>>& value() const { return _value; }
    void pushBack(const std::shared_ptr<T>& v) {
      _value.push_back(v);
      touch();
    }
    void fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>& resolve) override {
      if (v.kind != BossValue::Array) throw std::runtime_error("property " + _name + ": expected an array");
      _value.clear();
      for (const auto& i : v.items) {
        if (!i.isPointer()) throw std::runtime_error("property " + _name + ": expected #pointer items");
        ConfigurablePtr p = resolve(i.pointerId());
        std::shared_ptr<T> typed = std::dynamic_pointer_cast<T>(p);
        if (p && !typed) throw std::runtime_error("property " + _name + ": linked object has the wrong class");
        _value.push_back(typed);
      }
      touch();
    }
    void toText(std::ostream& os, const std::function<int(const Configurable*)>& idOf) const override {
      os << "[ ";
      for (size_t i = 0; i < _value.size(); ++i)
        os << (i ? ", " : "") << "{ \"#pointer\" : " << (_value[i] ? idOf((const Configurable*) _value[i].get()) : -1)
           << " }";
      os << " ]";
    }

  protected:
    std::vector<std::shared_ptr<T>> _value;
  };

// same shape as srrg2_core's macro: PARAM(PropertyType, name, "doc", default, &changed_flag_or_0)
#define PARAM(TYPE, NAME, DOC, DEFAULT, FLAG) TYPE param_##NAME = TYPE(#NAME, DOC, this, DEFAULT, FLAG)
#define PARAM_VECTOR(TYPE, NAME, DOC, FLAG) TYPE param_##NAME = TYPE(#NAME, DOC, this, FLAG)

  // ------------------------------------------------------------------ configurable + registry
  class Configurable {
  public:
    virtual ~Configurable() {}
    virtual std::string className() const { return _class_name; }
    const std::string& name() const { return _name; }
    void setName(const std::string& n) { _name = n; }
    PropertyBase* property(const std::string& name) const {
      auto it = _properties.find(name);
      return it == _properties.end() ? nullptr : it->second;
    }
    const std::map<std::string, PropertyBase*>& properties() const { return _properties; }
    // values of parameters this build does not model (kept so that a config survives read -> write)
    std::vector<std::pair<std::string, BossValue>> unknown_fields;

  protected:
    friend class PropertyBase;
    friend class ConfigurableManager;
    friend class ClassRegistry;
    std::map<std::string, PropertyBase*> _properties;
    std::string _name, _class_name;
  };

  // stands in for every class of a configuration that is outside the hot path (MultiGraphSLAM2D, ...)
  class GenericConfigurable : public Configurable {};

  class ClassRegistry {
  public:
    using Factory = std::function<ConfigurablePtr()>;
    static ClassRegistry& instance();
    void add(const std::string& class_name, Factory f) { _factories[class_name] = f; }
    bool has(const std::string& class_name) const { return _factories.count(class_name) != 0; }
    ConfigurablePtr create(const std::string& class_name) const;
    std::vector<std::string> classNames() const;

  private:
    std::map<std::string, Factory> _factories;
  };

#define BOSS_REGISTER_CLASS(CLASS) \
  srrg2_core::ClassRegistry::instance().add(#CLASS, []() { return srrg2_core::ConfigurablePtr(new CLASS); })
// register CLASS under the name a reference configuration uses for the module it replaces
#define BOSS_REGISTER_CLASS_AS(CLASS, NAME) \
  srrg2_core::ClassRegistry::instance().add(NAME, []() { return srrg2_core::ConfigurablePtr(new CLASS); })

  // ------------------------------------------------------------------ manager
  class ConfigurableManager {
  public:
    // instantiate every object of a BOSS-text configuration; unknown classes become GenericConfigurable
    void read(const std::string& filename);
    void readString(const std::string& text);
    void write(const std::string& filename) const;
    std::string writeString() const;

    template <typename T>
    std::shared_ptr<T> getByName(const std::string& name) const {
      for (const auto& o : _objects)
        if (o->name() == name) {
          if (auto t = std::dynamic_pointer_cast<T>(o)) return t;
        }
      return nullptr;
    }
    template <typename T>
    std::shared_ptr<T> getById(int id) const {
      auto it = _by_id.find(id);
      return it == _by_id.end() ? nullptr : std::dynamic_pointer_cast<T>(it->second);
    }
    template <typename T>
    std::vector<std::shared_ptr<T>> getAll() const {
      std::vector<std::shared_ptr<T>> out;
      for (const auto& o : _objects)
        if (auto t = std::dynamic_pointer_cast<T>(o)) out.push_back(t);
      return out;
    }
    template <typename T>
    std::shared_ptr<T> create(const std::string& name = "") {
      std::shared_ptr<T> t(new T);
      t->setName(name);
      add(t);
      return t;
    }
    void add(const ConfigurablePtr& o);
    const std::vector<ConfigurablePtr>& objects() const { return _objects; }
    int idOf(const Configurable* c) const;

  private:
    std::vector<ConfigurablePtr> _objects;
    std::map<int, ConfigurablePtr> _by_id;
    std::map<const Configurable*, int> _ids;
    int _next_id = 1;
  };

  // parser entry point (also used by the tests)
  std::vector<std::pair<std::string, BossValue>> parseBossText(const std::string& text);

}  // namespace srrg2_core
