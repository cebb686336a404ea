// ls2d_boss.cpp -- BOSS-text reader/writer, class registry, property plumbing (see ls2d_boss.h).
#include "ls2d_boss.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iomanip>

namespace srrg2_core {

  PropertyBase::PropertyBase(const std::string& name, const std::string& doc, Configurable* owner,
                             bool* changed_flag) :
    _name(name), _doc(doc), _changed_flag(changed_flag) {
    if (owner) owner->_properties[name] = this;
  }

  // ---- scalar properties <-> values
  template <>
  void Property_<float>::fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>&) {
    if (v.kind != BossValue::Number) throw std::runtime_error("property " + _name + ": expected a number");
    setValue((float) v.number);
  }
  template <>
  void Property_<int>::fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>&) {
    if (v.kind != BossValue::Number) throw std::runtime_error("property " + _name + ": expected a number");
    setValue((int) v.number);
  }
  template <>
  void Property_<unsigned>::fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>&) {
    if (v.kind != BossValue::Number) throw std::runtime_error("property " + _name + ": expected a number");
    setValue((unsigned) v.number);
  }
  template <>
  void Property_<bool>::fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>&) {
    if (v.kind != BossValue::Number) throw std::runtime_error("property " + _name + ": expected 0/1");
    setValue(v.number != 0);
  }
  template <>
  void Property_<std::string>::fromValue(const BossValue& v, const std::function<ConfigurablePtr(int)>&) {
    if (v.kind != BossValue::String) throw std::runtime_error("property " + _name + ": expected a string");
    setValue(v.text);
  }
  template <>
  void Property_<float>::toText(std::ostream& os, const std::function<int(const Configurable*)>&) const {
    os << std::setprecision(9) << _value;
  }
  template <>
  void Property_<int>::toText(std::ostream& os, const std::function<int(const Configurable*)>&) const {
    os << _value;
  }
  template <>
  void Property_<unsigned>::toText(std::ostream& os, const std::function<int(const Configurable*)>&) const {
    os << _value;
  }
  template <>
  void Property_<bool>::toText(std::ostream& os, const std::function<int(const Configurable*)>&) const {
    os << (_value ? 1 : 0);
  }
  template <>
  void Property_<std::string>::toText(std::ostream& os, const std::function<int(const Configurable*)>&) const {
    os << '"' << _value << '"';
  }

  // ---- registry
  ClassRegistry& ClassRegistry::instance() {
    static ClassRegistry r;
    return r;
  }
  ConfigurablePtr ClassRegistry::create(const std::string& class_name) const {
    auto it = _factories.find(class_name);
    if (it == _factories.end()) return nullptr;
    ConfigurablePtr p = it->second();
    p->_class_name    = class_name;  // keep the name the configuration used
    return p;
  }
  std::vector<std::string> ClassRegistry::classNames() const {
    std::vector<std::string> out;
    for (const auto& f : _factories) out.push_back(f.first);
    return out;
  }

  // ---- parser: a sequence of  "ClassName" { json-ish body }  with // comments
  namespace {
    struct Parser {
      const std::string& s;
      size_t i = 0;
      explicit Parser(const std::string& text) : s(text) {}
      [[noreturn]] void fail(const std::string& what) const {
        size_t line = 1;
        for (size_t k = 0; k < i && k < s.size(); ++k) line += s[k] == '\n';
        throw std::runtime_error("BOSS parse error at line " + std::to_string(line) + ": " + what);
      }
      void skip() {
        for (;;) {
          while (i < s.size() && std::isspace((unsigned char) s[i])) ++i;
          if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {
            while (i < s.size() && s[i] != '\n') ++i;
            continue;
          }
          break;
        }
      }
      bool eof() {
        skip();
        return i >= s.size();
      }
      char peek() {
        skip();
        if (i >= s.size()) fail("unexpected end of file");
        return s[i];
      }
      void expect(char c) {
        if (peek() != c) fail(std::string("expected '") + c + "'");
        ++i;
      }
      std::string string() {
        expect('"');
        std::string out;
        while (i < s.size() && s[i] != '"') {
          if (s[i] == '\\' && i + 1 < s.size()) ++i;
          out.push_back(s[i++]);
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
      }
      BossValue value() {
        BossValue v;
        const char c = peek();
        if (c == '"') {
          v.kind = BossValue::String;
          v.text = string();
        } else if (c == '{') {
          v = object();
        } else if (c == '[') {
          ++i;
          v.kind = BossValue::Array;
          if (peek() == ']') {
            ++i;
            return v;
          }
          for (;;) {
            v.items.push_back(value());
            if (peek() == ',') {
              ++i;
              continue;
            }
            expect(']');
            break;
          }
        } else if (c == '-' || c == '+' || c == '.' || std::isdigit((unsigned char) c)) {
          char* end = nullptr;
          v.kind    = BossValue::Number;
          v.number  = std::strtod(s.c_str() + i, &end);
          if (end == s.c_str() + i) fail("bad number");
          i = (size_t)(end - s.c_str());
        } else if (s.compare(i, 4, "null") == 0) {
          i += 4;
        } else if (s.compare(i, 4, "true") == 0) {
          i += 4;
          v.kind = BossValue::Number, v.number = 1;
        } else if (s.compare(i, 5, "false") == 0) {
          i += 5;
          v.kind = BossValue::Number, v.number = 0;
        } else {
          fail(std::string("unexpected character '") + c + "'");
        }
        return v;
      }
      BossValue object() {
        BossValue v;
        v.kind = BossValue::Object;
        expect('{');
        if (peek() == '}') {
          ++i;
          return v;
        }
        for (;;) {
          std::string key = string();
          expect(':');
          v.fields.emplace_back(key, value());
          if (peek() == ',') {
            ++i;
            continue;
          }
          expect('}');
          break;
        }
        return v;
      }
    };
  }  // namespace

  std::vector<std::pair<std::string, BossValue>> parseBossText(const std::string& text) {
    Parser p(text);
    std::vector<std::pair<std::string, BossValue>> out;
    while (!p.eof()) {
      std::string cls = p.string();
      out.emplace_back(cls, p.object());
    }
    return out;
  }

  // ---- manager
  void ConfigurableManager::add(const ConfigurablePtr& o) {
    if (_ids.count(o.get())) return;
    while (_by_id.count(_next_id)) ++_next_id;
    _ids[o.get()]  = _next_id;
    _by_id[_next_id] = o;
    _objects.push_back(o);
  }

  int ConfigurableManager::idOf(const Configurable* c) const {
    auto it = _ids.find(c);
    return it == _ids.end() ? -1 : it->second;
  }

  void ConfigurableManager::read(const std::string& filename) {
    std::ifstream is(filename);
    if (!is.good()) throw std::runtime_error("ConfigurableManager::read| cannot open " + filename);
    std::stringstream ss;
    ss << is.rdbuf();
    readString(ss.str());
  }

  void ConfigurableManager::readString(const std::string& text) {
    auto parsed = parseBossText(text);
    std::vector<ConfigurablePtr> created;
    // pass 1: instantiate by registered class name
    for (auto& po : parsed) {
      ConfigurablePtr obj = ClassRegistry::instance().create(po.first);
      if (!obj) {
        obj.reset(new GenericConfigurable);
        obj->_class_name = po.first;
      }
      const BossValue* id = po.second.find("#id");
      if (!id || id->kind != BossValue::Number) throw std::runtime_error("object \"" + po.first + "\" has no #id");
      const int oid = (int) id->number;
      if (_by_id.count(oid)) throw std::runtime_error("duplicate #id " + std::to_string(oid));
      if (const BossValue* nm = po.second.find("name")) obj->setName(nm->text);
      _by_id[oid]     = obj;
      _ids[obj.get()] = oid;
      _objects.push_back(obj);
      created.push_back(obj);
    }
    // pass 2: parameters and links
    auto resolve = [this](int id) -> ConfigurablePtr {
      if (id < 0) return nullptr;
      auto it = _by_id.find(id);
      if (it == _by_id.end()) throw std::runtime_error("dangling #pointer " + std::to_string(id));
      return it->second;
    };
    for (size_t k = 0; k < parsed.size(); ++k) {
      Configurable* obj = created[k].get();
      for (const auto& f : parsed[k].second.fields) {
        if (f.first == "#id" || f.first == "name") continue;
        PropertyBase* prop = obj->property(f.first);
        if (prop) {
          try {
            prop->fromValue(f.second, resolve);
          } catch (const std::runtime_error& e) {
            throw std::runtime_error("\"" + parsed[k].first + "\" #" + std::to_string(idOf(obj)) + ": " + e.what());
          }
        } else {
          obj->unknown_fields.push_back(f);
        }
      }
    }
  }

  namespace {
    void writeValue(std::ostream& os, const BossValue& v) {
      switch (v.kind) {
        case BossValue::Null: os << "null"; break;
        case BossValue::Number: os << std::setprecision(9) << v.number; break;
        case BossValue::String: os << '"' << v.text << '"'; break;
        case BossValue::Array:
          os << "[ ";
          for (size_t i = 0; i < v.items.size(); ++i) {
            if (i) os << ", ";
            writeValue(os, v.items[i]);
          }
          os << " ]";
          break;
        case BossValue::Object:
          os << "{ ";
          for (size_t i = 0; i < v.fields.size(); ++i) {
            if (i) os << ", ";
            os << '"' << v.fields[i].first << "\" : ";
            writeValue(os, v.fields[i].second);
          }
          os << " }";
          break;
      }
    }
  }  // namespace

  std::string ConfigurableManager::writeString() const {
    std::ostringstream os;
    auto id_of = [this](const Configurable* c) { return idOf(c); };
    for (const auto& o : _objects) {
      os << '"' << o->className() << "\" {\n  \"#id\" : " << idOf(o.get());
      if (!o->name().empty()) os << ",\n  \"name\" : \"" << o->name() << '"';
      for (const auto& p : o->properties()) {
        os << ",\n\n  // " << p.second->doc() << "\n  \"" << p.first << "\" : ";
        p.second->toText(os, id_of);
      }
      for (const auto& f : o->unknown_fields) {
        os << ",\n  \"" << f.first << "\" : ";
        writeValue(os, f.second);
      }
      os << "\n }\n\n";
    }
    return os.str();
  }

  void ConfigurableManager::write(const std::string& filename) const {
    std::ofstream os(filename);
    if (!os.good()) throw std::runtime_error("ConfigurableManager::write| cannot open " + filename);
    os << writeString();
  }

}  // namespace srrg2_core
